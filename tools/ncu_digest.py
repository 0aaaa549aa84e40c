"""Digest of an ncu report (read here, no GPU needed): headline counters per kernel + the source lines with the most
stall samples.   python tools/ncu_digest.py gpurun_out/x.ncu-rep [top_lines]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rr[2:]:
    print("###", r[hdr.index("Kernel Name")][:110])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"  {w:85s} {r[i]:>18s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# the source page prints one table per kernel, each starting with a header row containing "Source"
tables, cur = [], None
for r in rows:
    if "Source" in r and any("Sampl" in c for c in r):
        cur = {"hdr": r, "rows": []}; tables.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for t in tables:
    h = t["hdr"]; si = h.index("Source")
    samp = [i for i, c in enumerate(h) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Sampling Data (All)")]
    if not samp:
        samp = [i for i, c in enumerate(h) if "Sampl" in c][:1]
    k = samp[0]
    def val(r):
        try: return float(r[k].replace(",", ""))
        except ValueError: return 0.0
    tot = sum(val(r) for r in t["rows"]) or 1.0
    print(f"--- source table: {len(t['rows'])} lines, column '{h[k]}', total {tot:.0f}")
    for r in sorted(t["rows"], key=val, reverse=True)[:top]:
        print(f"  {100 * val(r) / tot:5.1f}%  {r[si][:150]}")
