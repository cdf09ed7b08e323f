"""Build variants of the library that differ in compile-time switches of one kernel, for A/B timing
on the GPU box in one call (each variant: QB200_LIB=<path> python tests/tools/prof_<kernel>.py).

    python tests/tools/kernel_variants.py            # builds qunundrum_b200/_variants/lib_<tag>.so

Only the one source file of a variant is recompiled; the other objects are the regular build's.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qunundrum_b200 import build as B  # noqa: E402

VARIANTS = {
    # tag: (source file, flags)
    "diagk_occ6": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=6"]),
    "diagk_occ8": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=8"]),
    "diagk_occ9": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=9"]),
    "diagk_occ10": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=10"]),
    "diagk_occ12": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=12"]),
    "diagk_occ8_unroll8": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=8", "-DQB_DIAGK_UNROLL=8"]),
    "diagk_occ8_unroll2": ("qb200_diagk.cu", ["-DQB_DIAGK_MIN_CTAS=8", "-DQB_DIAGK_UNROLL=2"]),
    "sample_occ6": ("qb200_sampler.cu", ["-DQB_SAMPLE_MIN_CTAS=6"]),
    "sample_occ7": ("qb200_sampler.cu", ["-DQB_SAMPLE_MIN_CTAS=7"]),
    "sample_occ8": ("qb200_sampler.cu", ["-DQB_SAMPLE_MIN_CTAS=8"]),
    "sample_occ10": ("qb200_sampler.cu", ["-DQB_SAMPLE_MIN_CTAS=10"]),
    "sample_occ12": ("qb200_sampler.cu", ["-DQB_SAMPLE_MIN_CTAS=12"]),
}


def main():
    B.build()
    out = os.path.join(os.path.dirname(B.LIB), "_variants")
    os.makedirs(out, exist_ok=True)
    objs = [os.path.join(B.OBJ, os.path.basename(s) + ".o") for s in B.sources()]
    procs = []
    for tag, (src, flags) in VARIANTS.items():
        obj = os.path.join(out, f"{tag}.o")
        cmd = [B.nvcc_path(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj]
        procs.append((tag, src, obj, subprocess.Popen(cmd)))
    for tag, src, obj, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"variant {tag} failed to compile")
        lib = os.path.join(out, f"lib_{tag}.so")
        link = [o if not o.endswith(src + ".o") else obj for o in objs]
        subprocess.check_call([B.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *link, "-o", lib])
        print(lib)


if __name__ == "__main__":
    main()
