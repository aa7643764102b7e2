#!/usr/bin/env python
"""Summarises an .ncu-rep (run in the build container: `ncu -i` needs no GPU).
usage: ncu_summary.py report.ncu-rep [--source N] [--json out.json KEY FRAMES]
--json merges {KEY: {metric: value, ..., "frames": FRAMES}} of the LAST kernel in the report into out.json
(bench.py reads roofline.traffic from profiles/r02_ncu_layered_i8.json)."""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for d in data:
    print("=" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {d[i]:>16s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
           h.startswith("smsp__average_warp_latency_issue_stalled_"):
            try:
                st.append((float(d[i]), h))
            except ValueError:
                pass
    print("-- warp stall reasons (warps per issue-active cycle)")
    for v, h in sorted(st, reverse=True)[:10]:
        print(f"   {v:8.3f} {h}")
if "--json" in sys.argv:
    j = sys.argv.index("--json")
    path, key, frames = sys.argv[j + 1], sys.argv[j + 2], int(sys.argv[j + 3])
    d = data[-1]
    rec = {"frames": frames, "report": os.path.basename(rep)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(d[i].replace(",", ""))
                if ("bytes" in k or "time_duration" in k) and units[i] in scale:
                    v *= scale[units[i]]     # bytes -> B, time -> ms
                rec[k] = v
            except ValueError:
                rec[k] = d[i]
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                rec[h] = float(d[i])
            except ValueError:
                pass
    allj = {}
    if os.path.exists(path):
        with open(path) as f:
            allj = json.load(f)
    allj[key] = rec
    with open(path, "w") as f:
        json.dump(allj, f, indent=1, sort_keys=True)
    print("wrote", path, key)
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(src)))
    h = rr[0]
    print("-- source page columns:", h)
