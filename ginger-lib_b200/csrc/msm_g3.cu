// group 3 of include/g753.h (see msm_impl.cuh)
#include "msm_impl.cuh"
G753_INSTANTIATE_GROUP(3)
