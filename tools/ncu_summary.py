"""Condense an .ncu-rep (ncu --set full) into one CSV row per captured launch with the metrics the roofline
discussion uses.   python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/xxx.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def main():
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel", "id"] + WANT)
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        ki, ii = hdr.index("Kernel Name"), hdr.index("ID")
        for r in rows[2:]:
            vals = []
            for m in WANT:
                vals.append(r[hdr.index(m)] + " " + rows[1][hdr.index(m)] if m in hdr else "")
            w.writerow([rep.split("/")[-1], r[ki].split("(")[0][:60], r[ii]] + vals)


if __name__ == "__main__":
    main()
