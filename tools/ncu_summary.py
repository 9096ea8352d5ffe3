#!/usr/bin/env python
"""Summarise an ncu report (no GPU needed): python tools/ncu_summary.py report.ncu-rep [metric-substring ...]
Prints, per profiled launch, the metrics the profiling guide names plus any whose name contains a given substring."""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_op_read_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
           "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')][:90]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if any(h.endswith(d) or h == d for d in DEFAULT) or any(e in h for e in extra):
                print(f"   {h} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
