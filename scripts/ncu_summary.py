"""One ncu report (--set full) -> the metrics the roofline discussion uses, as a small CSV under profiles/.
usage: python scripts/ncu_summary.py report.ncu-rep out.csv "header comment" """
import csv, subprocess, sys, io
rep, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum")
with open(dst, "w") as f:
    f.write(f"# {title}\n")
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.endswith(tuple(KEEP[1:])) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            f.write(f'{h},{v},{u}\n')
print(open(dst).read())
