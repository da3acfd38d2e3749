"""Summary of an `ncu --set full` report: one block per kernel launch with the metrics DESIGN.md quotes.
   python scripts/ncu_full_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_aperture_sysmem_op_write.sum", "pcie__write_bytes.sum", "pcie__read_bytes.sum")
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) <= ik: continue
    print("==", r[ik])
    for h, u, v in sorted(zip(hdr, units, r)):
        if h in keep or ("issue_stalled" in h and "per_issue_active" in h and v and float(v.replace(",", "")) > 0.3):
            print(f"  {h} [{u}] = {v}")
