"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none summary of {path}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\n== {name}")
        rd = wr = None
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:88s} {r[i]:>16s} {units[i]}")
                if w == "dram__bytes_read.sum":
                    rd = (float(r[i].replace(",", "")), units[i])
                if w == "dram__bytes_write.sum":
                    wr = (float(r[i].replace(",", "")), units[i])
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1)
            t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            tu = units[hdr.index("gpu__time_duration.sum")]
            ts = t * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(tu, 1e-3)
            print(f"  {'traffic = dram read+write':88s} {tot / 1e9:16.3f} GB  ({tot / ts / 1e9:.0f} GB/s under ncu)")


if __name__ == "__main__":
    main(sys.argv[1])
