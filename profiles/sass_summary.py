#!/usr/bin/env python
"""SASS opcode histograms of the product kernels (libb200conv.so), one text file per kernel under profiles/.

    python profiles/sass_summary.py            # writes profiles/r02_sass_<kernel>.txt

What to look for: UTCHMMA / UTCBAR / LDTM (tcgen05.mma / commit / ld: the tensor-core FIR), UBLKCP (1-D TMA bulk
copies) and SYNCS (mbarrier) in the FIR and tensor-core kernels, FFMA density in fir_direct, LDG.E.128 streams in
the UPOLS kernels; no library kernels (cuFFT / cuBLAS) anywhere in the image.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpuaudiobench_b200", "lib", "libb200conv.so")
KERNELS = {
    "tc_toeplitz": r"tc_toeplitz_kernel",
    "fir_direct_A32": r"fir_direct_kernelILi32E",
    "fir_direct_A4": r"fir_direct_kernelILi4E",
    "upols_fused_4x4": r"upols_fused_kernelILi4ELi4ELb0ELb0E",
    "upols_fused_4x4_strip": r"upols_fused_kernelILi4ELi4ELb1ELb0E",
    "fdl_mac": r"fdl_mac_kernel",
    "rfft_fwd": r"rfft_fwd_kernel",
    "irfft_ols": r"irfft_ols_kernel",
    "mix_rows": r"mix_rows_kernel",
    "mix_cluster": r"mix_cluster_kernel",
    "bus_allreduce": r"bus_allreduce_kernel",
    "strip": r"strip_",
}
NOTABLE = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "UCGABAR", "PREEXIT",
           "ACQBULK", "FFMA", "FFMA2", "LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMG", "MEMBAR", "BAR", "NANOSLEEP", "CCTL")


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", text)))
    for name, pat in KERNELS.items():
        matches = [f for f in funcs[1:] if re.search(pat, f.split("\n", 1)[0])]
        if not matches:
            continue
        out = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — opcode histogram; cubin architectures in the library: {', '.join(archs)}"]
        for f in matches:
            head = f.split("\n", 1)[0].strip()
            ops = collections.Counter()
            mods = collections.Counter()
            for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", f, re.M):
                ops[m.group(1)] += 1
                if m.group(1) in ("LDG", "STG", "LDS", "STS", "UTCHMMA", "LDTM", "UBLKCP", "SYNCS", "ATOMG", "MEMBAR"):
                    mods[m.group(1) + m.group(2)] += 1
            total = sum(ops.values())
            out.append(f"\n== {head}\ninstructions: {total}")
            out.append("notable: " + ", ".join(f"{k} {ops[k]}" for k in NOTABLE if ops.get(k)))
            out.append("top opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
            out.append("memory / async forms: " + ", ".join(f"{k} {v}" for k, v in mods.most_common(16)))
        path = os.path.join(ROOT, "profiles", f"r02_sass_{name}.txt")
        with open(path, "w") as fh:
            fh.write("\n".join(out) + "\n")
        print("wrote", os.path.relpath(path, ROOT))


if __name__ == "__main__":
    sys.exit(main())
