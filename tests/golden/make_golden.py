#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors and constants into JSON fixtures.

Runs ONLY in the build container (needs /root/reference, which does not exist on
the GPU box).  The output files next to this script are committed, so the tests
never touch /root/reference at run time.

Every 768-bit literal in the listed reference files is a Rust array
`BigInteger768([l0, ..., l11])` (or the alias `BigInteger([...])`) with l0 the
least-significant u64 (algebra/src/biginteger/mod.rs:20-27).  We record, in
source order, each literal with
  * the enclosing `fn test_*` / `const NAME` scope,
  * the wrapper that says how to read the limbs:
      "mont"  -> Fq::new(..) / field_new!(..): raw Montgomery limbs
      "canon" -> Fq::from_repr(..) / Fr::from_repr(..): canonical integer
      "raw"   -> a bare BigInteger (MODULUS, R, scalars ...).
The tests (tests/test_oracle_kat.py) know the shape of each reference test and
replay it against the oracle.

usage: python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

REF = "/root/reference/algebra/src"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "fields_mnt4753_tests": "fields/mnt4753/tests.rs",
    "fields_mnt6753_tests": "fields/mnt6753/tests.rs",
    "curves_mnt4753_tests": "curves/mnt4753/tests.rs",
    "curves_mnt6753_tests": "curves/mnt6753/tests.rs",
    "fields_mnt4753_fq": "fields/mnt4753/fq.rs",
    "fields_mnt6753_fq": "fields/mnt6753/fq.rs",
    "fields_mnt4753_fq2": "fields/mnt4753/fq2.rs",
    "fields_mnt6753_fq3": "fields/mnt6753/fq3.rs",
    "curves_mnt4753_g1": "curves/mnt4753/g1.rs",
    "curves_mnt4753_g2": "curves/mnt4753/g2.rs",
    "curves_mnt4753_mod": "curves/mnt4753/mod.rs",
    "curves_mnt6753_g1": "curves/mnt6753/g1.rs",
    "curves_mnt6753_g2": "curves/mnt6753/g2.rs",
    "curves_mnt6753_mod": "curves/mnt6753/mod.rs",
}

# only keep the scopes the hot path needs (keeps the fixtures small)
KEEP_TEST_SCOPES = re.compile(
    r"test_fq_(add_assign|sub_assign|mul_assign|squaring|root_of_unity)$|"
    r"test_fq[23]_(squaring|mul|inverse|addition|subtraction|negation|doubling|mul_nonresidue)$|"
    r"test_g[12]_(addition_correctness|doubling_correctness|scalar_multiplication|affine_projective_conversion)$"
)

LIT = re.compile(r"BigInteger(?:768)?\(\s*\[(.*?)\]\s*\)", re.S)
SMALL = re.compile(r"BigInteger(?:768)?::from\((\d+)\)")
SCOPE = re.compile(r"^\s*(?:pub\s+)?(?:fn\s+(\w+)|const\s+(\w+)\s*:)", re.M)


def wrapper_before(text, pos):
    head = text[max(0, pos - 40):pos]
    head = head.rstrip()
    if re.search(r"from_repr\($", head):
        return "canon"
    if re.search(r"(Fq|Fr)::new\($", head) or re.search(r"field_new!\(\s*(Fq|Fr)\s*,$", head):
        return "mont"
    return "raw"


def extract(path):
    text = open(path).read()
    scopes = [(m.start(), m.group(1) or ("const:" + m.group(2))) for m in SCOPE.finditer(text)]
    items = []
    events = []
    for m in LIT.finditer(text):
        # strip comments inside the literal (e.g. "// = COEFF_A") before splitting
        body = re.sub(r"//[^\n]*", "", m.group(1))
        toks = [t.strip() for t in body.replace("\n", " ").split(",") if t.strip()]
        limbs = [int(t.replace("_", ""), 0) for t in toks]
        if len(limbs) != 12:
            continue
        val = sum(l << (64 * i) for i, l in enumerate(limbs))
        events.append((m.start(), wrapper_before(text, m.start()), val))
    for m in SMALL.finditer(text):
        events.append((m.start(), wrapper_before(text, m.start()), int(m.group(1))))
    events.sort()
    for pos, wrap, val in events:
        scope = None
        for s_pos, name in scopes:
            if s_pos <= pos:
                scope = name
            else:
                break
        line = text.count("\n", 0, pos) + 1
        items.append({"scope": scope, "line": line, "wrap": wrap, "value": hex(val)})
    return items


def small_consts(path):
    """TWO_ADICITY / INV / MODULUS_BITS style scalar constants."""
    text = open(path).read()
    out = {}
    for m in re.finditer(r"const\s+(\w+)\s*:\s*u(?:32|64)\s*=\s*([0-9a-fA-Fx_]+)\s*;", text):
        out[m.group(1)] = int(m.group(2).replace("_", ""), 0)
    return out


def pairing_consts(path):
    """the pairing parameters of curves/mnt{4,6}753/mod.rs that are not 768-bit literals: the Miller loop
    count (&[u64]), its signed-digit form WNAF (&[i32]) and the sign flags"""
    text = open(path).read()
    out = {}
    m = re.search(r"const\s+ATE_LOOP_COUNT\s*:[^=]*=\s*&\[(.*?)\];", text, re.S)
    limbs = [int(t.strip(), 0) for t in m.group(1).split(",") if t.strip()]
    out["ATE_LOOP_COUNT"] = hex(sum(l << (64 * i) for i, l in enumerate(limbs)))
    m = re.search(r"const\s+WNAF\s*:[^=]*=\s*&\[(.*?)\];", text, re.S)
    out["WNAF"] = [int(t.strip()) for t in m.group(1).split(",") if t.strip()]
    for name in ("ATE_IS_LOOP_COUNT_NEG", "FINAL_EXPONENT_LAST_CHUNK_W0_IS_NEG"):
        m = re.search(r"const\s+%s\s*:\s*bool\s*=\s*(true|false)\s*;" % name, text)
        out[name] = m.group(1) == "true"
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    kat = {}
    params = {}
    for key, rel in FILES.items():
        path = os.path.join(REF, rel)
        items = extract(path)
        if key.endswith("_tests"):
            by_scope = {}
            for it in items:
                if it["scope"] and KEEP_TEST_SCOPES.search(it["scope"]):
                    by_scope.setdefault(it["scope"], []).append(
                        {"wrap": it["wrap"], "value": it["value"], "line": it["line"]})
            kat[key] = {"source": "algebra/src/" + rel, "tests": by_scope}
        else:
            by_const = {}
            for it in items:
                if it["scope"] and it["scope"].startswith("const:"):
                    by_const.setdefault(it["scope"][6:], []).append(
                        {"wrap": it["wrap"], "value": it["value"], "line": it["line"]})
            params[key] = {"source": "algebra/src/" + rel, "consts": by_const,
                           "ints": small_consts(path)}
    with open(os.path.join(HERE, "reference_kat.json"), "w") as f:
        json.dump(kat, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, "reference_params.json"), "w") as f:
        json.dump(params, f, indent=0, sort_keys=True)
    pairing = {key: dict(pairing_consts(os.path.join(REF, FILES[key])), source="algebra/src/" + FILES[key])
               for key in ("curves_mnt4753_mod", "curves_mnt6753_mod")}
    with open(os.path.join(HERE, "reference_pairing.json"), "w") as f:
        json.dump(pairing, f, indent=0, sort_keys=True)
    # the 96-byte serialisation fixtures (fields/mnt{4,6}753/test_vec/*_tobyte)
    for name in ("mnt4753", "mnt6753"):
        p = os.path.join(REF, "fields", name, "test_vec", name + "_tobyte")
        if os.path.exists(p):
            with open(p, "rb") as f:
                data = f.read()
            with open(os.path.join(HERE, name + "_tobyte.hex"), "w") as f:
                f.write(data.hex() + "\n")
    n_kat = sum(len(v) for d in kat.values() for v in d["tests"].values())
    n_par = sum(len(v) for d in params.values() for v in d["consts"].values())
    print("wrote %d KAT literals, %d parameter literals" % (n_kat, n_par))


if __name__ == "__main__":
    main()
