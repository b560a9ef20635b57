"""Compact text summary of an ncu report (--set full): per kernel launch the metrics DESIGN.md and
bench.py's roofline line refer to.  usage: ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_atom.sum"]
stalls = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
print(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (one block per profiled launch)")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"\n== {name[:100]}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]:>16s} {units[i]}")
    try:
        rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}      # ncu picks a unit per column
        tot = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
        print(f"  {'traffic = dram read + write':70s} {tot / 1e9:16.6f} Gbyte")
    except Exception:
        pass
    top = sorted(((float(r[hdr.index(h)] or 0), h) for h in stalls), reverse=True)[:6]
    print("  top stall reasons (warps per issue-active cycle): " +
          ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in top))
