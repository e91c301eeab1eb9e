"""Summarise an .ncu-rep (read here without a GPU): one line per captured launch with the judged metrics."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "xu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu%"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%")]
for r in rows[2:]:
    name = r[col["Kernel Name"]][:70]
    parts = []
    for k, short in want:
        if k in col:
            parts.append(f"{short}={r[col[k]]}{units[col[k]] if short in ('dur', 'dram_rd', 'dram_wr') else ''}")
    print(name, "|", " ".join(parts))
