"""Summarise an .ncu-rep (ncu --set full) into a small CSV/markdown table for profiles/: one row per captured launch.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_xxx_ncu_summary.csv"""
import csv
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time_us"), ("sm__cycles_elapsed.avg", "sm_cycles"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct")]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    cols = [(k, n) for k, n in WANT if k in idx]
    w.writerow([n + ("[%s]" % units[idx[k]] if units[idx[k]] else "") for k, n in cols])
    for r in rows[2:]:
        w.writerow([r[idx[k]].replace("void casync::<unnamed>::", "").replace("casync::<unnamed>::", "") for k, _ in cols])


if __name__ == "__main__":
    main(sys.argv[1])
