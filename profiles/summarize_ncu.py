#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into the small text/JSON files committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_rats_v2.ncu-rep profiles/r1_rats_plate_v2 [--chains 1024 --workload rats]
"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_ldgsts.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    extra = dict(zip(sys.argv[3::2], sys.argv[4::2]))
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    rec = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h in ("Kernel Name",):
            rec[h] = f"{v} {u}".strip()
    src = page(rep, "source")
    sh = src[1]
    i_src, i_ex = sh.index("Source"), sh.index("Instructions Executed")
    rows = [(r[i_src], int(r[i_ex] or 0)) for r in src[2:] if len(r) > i_ex]
    total = sum(x[1] for x in rows)
    mix = Counter()
    for s, n in rows:
        parts = s.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        mix[op.split(".")[0]] += n
    with open(dst + "_summary.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
        for k in ["Kernel Name"] + KEYS:
            if k in rec:
                f.write(f"{k}: {rec[k]}\n")
        f.write(f"\n# executed warp-instruction mix (top 16 of {total}):\n")
        for op, n in mix.most_common(16):
            f.write(f"{op:10s} {n:14d} {100.0 * n / total:5.1f}%\n")
    def num(k):
        return float(rec[k].split()[0])
    unit = lambda k: rec[k].split()[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    tr = num("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + \
        num("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    if "--workload" in extra:
        tpath = "profiles/traffic.json"
        try:
            t = json.load(open(tpath))
        except Exception:
            t = {}
        t[extra["--workload"]] = {"dram_bytes_per_launch": tr, "chains": int(extra.get("--chains", 0)),
                                  "source": dst + "_summary.txt"}
        json.dump(t, open(tpath, "w"), indent=1)
    print(open(dst + "_summary.txt").read())


if __name__ == "__main__":
    main()
