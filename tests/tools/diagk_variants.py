"""Build variants of the library that differ in compile-time switches of k_diagk, for A/B timing
on the GPU box in one call (each variant: QB200_LIB=<path> python tests/tools/prof_diagk.py).

    python tests/tools/diagk_variants.py            # builds qunundrum_b200/_variants/lib_<tag>.so

Only qb200_diagk.cu is recompiled per variant; the other objects are the regular build's.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qunundrum_b200 import build as B  # noqa: E402

VARIANTS = {
    "base": [],
    "unroll8": ["-DQB_DIAGK_UNROLL=8"],
    "lds": ["-DQB_DIAGK_LDS=1"],
    "lds_unroll8": ["-DQB_DIAGK_LDS=1", "-DQB_DIAGK_UNROLL=8"],
    "lds_unroll8_occ5": ["-DQB_DIAGK_LDS=1", "-DQB_DIAGK_UNROLL=8", "-DQB_DIAGK_MIN_CTAS=5"],
    "lds_occ8": ["-DQB_DIAGK_LDS=1", "-DQB_DIAGK_MIN_CTAS=8"],
}


def main():
    B.build()
    out = os.path.join(os.path.dirname(B.LIB), "_variants")
    os.makedirs(out, exist_ok=True)
    objs = [os.path.join(B.OBJ, os.path.basename(s) + ".o") for s in B.sources()]
    src = os.path.join(B.CSRC, "qb200_diagk.cu")
    procs = []
    for tag, flags in VARIANTS.items():
        obj = os.path.join(out, f"qb200_diagk_{tag}.o")
        procs.append((tag, obj, subprocess.Popen([B.nvcc_path(), *B.NVCC_FLAGS, *flags, "-c", src, "-o", obj])))
    for tag, obj, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"variant {tag} failed to compile")
        lib = os.path.join(out, f"lib_{tag}.so")
        link = [o if not o.endswith("qb200_diagk.cu.o") else obj for o in objs]
        subprocess.check_call([B.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *link, "-o", lib])
        print(lib)


if __name__ == "__main__":
    main()
