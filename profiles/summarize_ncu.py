#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_c2.csv  > profiles/r01_c2_launches.txt
    python profiles/summarize_ncu.py full     gpurun_out/fir.ncu-rep      > profiles/r01_fir_direct_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) > vi:
            per.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    unit = rows[start + 1][ui]
    total = sum(sum(v) for k, v in per.items() if "b200conv" in k)
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {path}")
    print(f"# share = kernel's share of all b200conv kernel time in the capture")
    print(f"{'kernel':70s} {'launches':>8s} {'mean_' + unit:>12s} {'min_' + unit:>12s} {'share':>7s}")
    for k, v in per.items():
        share = f"{100 * sum(v) / total:6.1f}%" if "b200conv" in k else "      -"
        print(f"{k[:70]:70s} {len(v):8d} {sum(v) / len(v):12.1f} {min(v):12.1f} {share}")


def full(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on : {rep}")
    for n, r in enumerate(rows[2:]):
        print(f"\n== launch {n}: {r[hdr.index('Kernel Name')]}")
        for k in RAW_KEYS:
            if k in hdr:
                print(f"{k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    src = ncu_csv(rep, "source")
    shdr = src[1]
    body = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        body.append(r)
    si, ie = shdr.index("# Samples"), shdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in body) or 1
    print("\n== warp-state samples, first launch (source page)")
    agg = {shdr[i]: sum(int(r[i]) for r in body) for i in stall_cols}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"{k:28s} {v:8d} {100 * v / tot:5.1f}%")
    byop = collections.Counter()
    for r in body:
        toks = r[1].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += int(r[ie])
    te = sum(byop.values()) or 1
    for op in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS"):
        if byop.get(op) and op not in dict(byop.most_common(14)):
            print(f"(also executed: {op} {byop[op]})")
    print("\n== executed warp instructions by opcode, first launch")
    for op, c in byop.most_common(14):
        print(f"{op:10s} {c:12d} {100 * c / te:5.1f}%")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
