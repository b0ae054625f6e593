"""Summarise an .ncu-rep (raw page + source page) into a short text: key throughput metrics and the SASS lines
with the most stall samples.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
for r in rows[2:]:
    print("== kernel:", r[idx["Kernel Name"]][:80])
    for w in want:
        if w in idx:
            print(f"  {w} = {r[idx[w]]} {units[idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page holds one table per kernel; take the last one
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for st in starts[-1:]:
    hdr = rows[st]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[st + 1:] if len(r) == len(hdr)]
    f = lambda x: float(x) if x not in ("", None) else 0.0
    tot = sum(f(r[idx["# Samples"]]) for r in data) or 1.0
    toti = sum(f(r[idx["Instructions Executed"]]) for r in data) or 1.0
    print(f"-- source page: {int(tot)} samples, {int(toti)} warp instructions; top lines by samples")
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for r in sorted(data, key=lambda r: -f(r[idx["# Samples"]]))[:nl]:
        st_ = sorted(((k, f(r[idx[k]])) for k in keys), key=lambda kv: -kv[1])[:2]
        print(f"  {f(r[idx['# Samples']]) / tot * 100:5.1f}%  exec {f(r[idx['Instructions Executed']]) / toti * 100:5.2f}%  {r[idx['Source']][:70]:70s} {st_[0][0]}={int(st_[0][1])} {st_[1][0]}={int(st_[1][1])}")
