"""Static SASS opcode histogram of the hot kernels (no GPU needed):

    python tests/tools/sass_histogram.py [profiles/r02_sass_histogram]

cuobjdump -sass on the built library; per kernel (k_fused2d class variants, k_fused1d, k_diagk,
k_sample, k_text_format, k_collapse, ...): total instructions and the FP64 / integer-multiply /
memory mnemonics. Static counts (the whole kernel, every path once), the evidence next to the
executed counts of the ncu captures: what the FP64 pipe is fed with (DFMA vs DMUL vs DADD), that no
tensor-core / TMA instruction is present (there is no contraction to feed them), and how
large the kernels are."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "qunundrum_b200", "libqunundrum_b200.so")
WANT = ("k_fused2d", "k_fused1d", "k_fused_cols", "k_axis2d", "k_diagk", "k_sample", "k_text_format",
        "k_text_parse", "k_collapse", "k_scale_x87", "k_so_step", "k_vals1d", "k_dfma_peak")
GROUPS = {
    "fp64": ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"),
    "fp64_convert": ("F2F", "I2F", "F2I", "FRND"),
    "sfu": ("MUFU",),
    "imad": ("IMAD",),
    "int": ("IADD3", "LOP3", "SHF", "LEA", "ISETP", "SEL", "PRMT", "FLO", "POPC", "IABS", "BREV"),
    "shuffle_vote": ("SHFL", "VOTE", "MATCH", "REDUX"),
    "global_mem": ("LDG", "STG", "LDC", "LD", "ST", "ATOM", "ATOMG", "RED"),
    "shared_mem": ("LDS", "STS", "LDSM", "ATOMS"),
    "local_mem": ("LDL", "STL"),
    "async_copy": ("LDGSTS", "LDGDEPBAR", "DEPBAR", "UBLKCP", "UTMALDG", "UTMASTG"),
    "tensor": ("HMMA", "IMMA", "DMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "TCGEN"),
    "branch": ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "BAR"),
}


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.split("\n")
    rows = []
    for (mangled, hist), name in zip(kernels.items(), demangled):
        if not any(w in name for w in WANT):
            continue
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("qb200::", "")
        row = {"kernel": short, "instructions": sum(hist.values())}
        for g, ops in GROUPS.items():
            row[g] = {op: hist[op] for op in ops if hist[op]}
        row["fp64_total"] = sum(hist[o] for o in GROUPS["fp64"])
        row["dfma_share_of_fp64"] = (hist["DFMA"] / row["fp64_total"]) if row["fp64_total"] else None
        rows.append(row)
    rows.sort(key=lambda r: r["kernel"])
    json.dump(rows, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        f.write("static SASS histogram of libqunundrum_b200.so (cuobjdump -sass, sm_100a); whole kernels, every path once\n")
        f.write(f"{'kernel':58s} {'inst':>7s} {'DFMA':>6s} {'DMUL':>6s} {'DADD':>6s} {'DSETP':>6s} {'MUFU':>5s} {'IMAD':>6s} "
                f"{'LDG/STG':>8s} {'LDS/STS':>8s} {'LDL/STL':>8s} {'tensor':>6s}\n")
        for r in rows:
            g = lambda grp, op: r[grp].get(op, 0)  # noqa: E731
            f.write(f"{r['kernel'][:58]:58s} {r['instructions']:7d} {g('fp64', 'DFMA'):6d} {g('fp64', 'DMUL'):6d} "
                    f"{g('fp64', 'DADD'):6d} {g('fp64', 'DSETP'):6d} {g('sfu', 'MUFU'):5d} {g('imad', 'IMAD'):6d} "
                    f"{g('global_mem', 'LDG') + g('global_mem', 'STG'):8d} {g('shared_mem', 'LDS') + g('shared_mem', 'STS'):8d} "
                    f"{g('local_mem', 'LDL') + g('local_mem', 'STL'):8d} {sum(r['tensor'].values()):6d}\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
