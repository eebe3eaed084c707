"""Prints the key metrics of an .ncu-rep (raw page) and the hottest source lines (source page)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "launch__shared_mem_per_block_dynamic"]
for r in rows[2:]:
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k:
                print(f"{k:72s} {r[i]:>18s} {units[i]}")
    print("-- stall reasons (per warp active, %):")
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"   {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):32s} {v:8.2f}")
if len(sys.argv) > 2:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    try:
        ci = hdr.index("Instructions Executed"); si = hdr.index("Source"); wi = hdr.index("Warp Stall Sampling (All Samples)")
    except ValueError:
        print(hdr); sys.exit()
    tot = sum(int(r[ci] or 0) for r in rows[1:] if len(r) > ci)
    tots = sum(int(r[wi] or 0) for r in rows[1:] if len(r) > wi)
    print("total instr", tot, "samples", tots)
    top = sorted(rows[1:], key=lambda r: -int(r[wi] or 0))[:int(sys.argv[2])]
    for r in top:
        print(f"{int(r[ci] or 0)/tot*100:6.2f}% inst {int(r[wi] or 0)/max(1,tots)*100:6.2f}% stall  {r[si][:110]}")
