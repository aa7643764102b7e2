#!/usr/bin/env python
"""Imports the reference's INPUT DATA files (code definitions, not source code) into
cuda_ldpc_b200/data/ so that tests and bench.py can run on the GPU box, where /root/reference
does not exist.  Whitespace is normalised (LF line ends, single spaces); the numbers are
untouched and the on-disk formats are the reference's, which the loaders
keep verbatim.  GF arithmetic tables are NOT imported: they are regenerated from the primitive
polynomial (cuda_ldpc_b200/gf.py) and only their sha256 is pinned (tests/golden/nb_ref.json).

  python tools/import_reference_data.py          # build container only
"""
import glob, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF_ROOT", "/root/reference")
B = os.path.join(REF, "bldpc_实习")
NB = os.path.join(REF, "myNBLDPC")


def norm(src, dst):
    with open(src, "rb") as f:
        txt = f.read().decode("utf-8", "replace").replace("\r", "")
    lines = [re.sub(r"[ \t]+", " ", ln).strip() for ln in txt.split("\n")]
    while lines and not lines[-1]:
        lines.pop()
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    for p in sorted(glob.glob(os.path.join(B, "*_BlockH.txt"))) + [os.path.join(B, "PON_LDPC.txt")]:
        norm(p, os.path.join(ROOT, "cuda_ldpc_b200", "data", "bldpc", os.path.basename(p)))
    for name in ("BDS.576.288.GF.64.txt", "LDPC_N576_K288_GF64_d1_exp.txt", "LDPC_N576_K480_GF256_exp.txt",
                 "LDPC_N96_K48_GF256_d1_exp.txt", "Tanner_74_9_Z128_GF16.txt"):
        norm(os.path.join(NB, name), os.path.join(ROOT, "cuda_ldpc_b200", "data", "nbldpc", name))
    for p in sorted(glob.glob(os.path.join(NB, "Constellation", "*.txt"))):
        norm(p, os.path.join(ROOT, "cuda_ldpc_b200", "data", "nbldpc", "Constellation", os.path.basename(p)))


if __name__ == "__main__":
    main()
