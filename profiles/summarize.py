"""profiles/summarize.py <tag> -- condense the ncu exports that profiles/ncu_capture.sh left in
gpurun_out/ into the small, tracked files under profiles/ (the .ncu-rep binaries stay out of git)."""
import csv
import json
import os
import re
import shutil
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
N_PARTICLES = float(sys.argv[2]) if len(sys.argv) > 2 else 1e7      # particles of the profiled run (bench.py: 1e7)
SRC, DST = "gpurun_out", "profiles"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def raw(fn):
    rows = list(csv.reader(open(fn)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1], rows[i + 2]
    raise SystemExit(f"no kernel row in {fn}")


def mix(fn):
    rows = list(csv.reader(open(fn)))
    for i, r in enumerate(rows):
        if "Source" in r and "Address" in r:
            hdr, start = r, i + 1
            break
    iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
    tot, ops = 0.0, {}
    for r in rows[start:]:
        try:
            e = float(r[iE])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] = ops.get(op, 0) + e
        tot += e
    return tot, dict(sorted(ops.items(), key=lambda kv: -kv[1])[:24])


out = {}
for k, label in [("move", "step kernel, no-resampling step (config 2's common case)"),
                 ("movers", "step kernel, resampling step (weights -> CDF, grid barrier, search + gather + move)"),
                 ("scan", "k_scan_w, resampling step (weights -> CDF; round-1 kernels only)")]:
    fn = f"{SRC}/{k}_{TAG}_raw.csv"
    if not os.path.exists(fn):
        continue
    hdr, units, data = raw(fn)
    d = {"what": label, "kernel": data[hdr.index("Kernel Name")]}
    for key in KEYS:
        if key in hdr:
            j = hdr.index(key)
            try:
                d[key] = [float(data[j]), units[j]]
            except ValueError:
                pass
    tot, ops = mix(f"{SRC}/{k}_{TAG}_source.csv")
    d["warp_instructions_executed"] = tot
    d["top_opcodes"] = ops
    # thread-level instructions per PAIR of particles (the kernel's work unit): warp instructions / (N / 2 / 32)
    pairs_warp = N_PARTICLES / 2 / 32
    d["instructions_per_pair"] = tot / pairs_warp
    d["fp64_inst_per_pair"] = sum(v for op, v in ops.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")) / pairs_warp
    rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
    if rd and wr:
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        d["dram_traffic_bytes"] = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
    out[k] = d
    for ext in ("details",):
        shutil.copy(f"{SRC}/{k}_{TAG}_{ext}.csv", f"{DST}/{TAG}_{k}_{ext}.csv")
json.dump(out, open(f"{DST}/{TAG}_ncu_summary.json", "w"), indent=1)
if os.path.exists(f"{SRC}/launches_{TAG}.csv"):
    shutil.copy(f"{SRC}/launches_{TAG}.csv", f"{DST}/{TAG}_launches.csv")
for k, d in out.items():
    print(k, d["kernel"][:50], {kk: vv for kk, vv in d.items() if kk in ("gpu__time_duration.sum", "dram_traffic_bytes")})
