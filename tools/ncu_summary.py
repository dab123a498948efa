"""Write a short text summary (key raw metrics + top stall lines) of an .ncu-rep capture.
usage: ncu_summary.py in.ncu-rep out.txt [launch_index]"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
launch = sys.argv[3] if len(sys.argv) > 3 else "0"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(vals, units)))
lines = [f"source: {rep} (launch {launch}); ncu --set full --clock-control none", f"kernel: {d.get('Kernel Name', ('?',))[0]}", ""]
for k in KEYS:
    if k in d:
        lines.append(f"{k:85s} {d[k][0]:>16s} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
try:
    h = srows[1]
    iS, iA = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    stall = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    data = []
    for r in srows[2:]:
        try:
            data.append((int(r[iA]), r))
        except Exception:
            pass
    tot = sum(x[0] for x in data) or 1
    agg = {x: 0 for x in stall}
    for n, r in data:
        for x in stall:
            try:
                agg[x] += int(r[h.index(x)])
            except Exception:
                pass
    lines += ["", f"warp stall samples: {tot}", "by reason: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]),
              "", "top SASS lines by stall samples:"]
    for n, r in sorted(data, key=lambda x: -x[0])[:16]:
        lines.append(f"  {100 * n / tot:5.1f}%  {r[iS].strip()[:100]}")
except Exception as e:
    lines.append(f"(no source page: {e})")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
