"""Condense ncu --set full reports into the tracked summaries under profiles/.
usage: ncu_profile_csv.py out.csv out_traffic.json cfg=report.ncu-rep [cfg=report.ncu-rep ...]"""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg.per_second"]
out_csv, out_json = sys.argv[1], sys.argv[2]
traffic = {"how": "ncu --set full --clock-control none --import-source on -k regex:spz_(en|de)code_(lane|wide) -s 2 -c 2 python bench.py "
                  "--config <c> --no-sub --no-cpu-baseline --steps 1 --warmup 3 (full BASELINE sizes, tools/final_run.sh); per launch",
           "configs": {}}
with open(out_csv, "w") as f:
    for arg in sys.argv[3:]:
        cfg, rep = arg.split("=")
        chunk_rows = 512
        if ":" in cfg:
            cfg, cr = cfg.split(":"); chunk_rows = int(cr)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = name.split("<")[0].replace("void ", "")
            f.write(f"## {cfg} {short}\nKernel Name,{name},\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"{k},{r[hdr.index(k)]},{units[hdr.index(k)]}\n")
            st = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
            for v, h in sorted(st, reverse=True)[:7]:
                f.write(f"stall_{h.split('issue_stalled_')[1].split('_per_')[0]}_per_issue,{v:.3f},ratio\n")
            f.write("\n")
            def val(k):
                v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            du, dv = units[hdr.index("gpu__time_duration.sum")], float(r[hdr.index("gpu__time_duration.sum")])
            traffic["configs"].setdefault(cfg, {"chunk_rows": chunk_rows})[short] = {
                "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
                "ncu_duration_us": dv * {"ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}[du], "kernel": name}
with open(out_json, "w") as f:
    json.dump(traffic, f, indent=1)
print(open(out_csv).read()[:3000])
