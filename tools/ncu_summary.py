"""Summarise an .ncu-rep (ncu --set full capture) per kernel launch: duration, DRAM bytes and
throughput, L2 hit rate, active warps, registers, pipe utilisation.  Runs where ncu is installed
(no GPU needed):  python tools/ncu_summary.py profiles/foo.ncu-rep > profiles/foo_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "active warps % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 inst % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA sub-pipe active %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall: long scoreboard (ratio)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {name: k for k, name in enumerate(head)}
    print(f"# {path}: {len(data)} profiled launch(es); ncu --set full --clock-control none")
    for r in data:
        print(f"\n{r[col['Kernel Name']][:150]}")
        print(f"  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        seen = set()
        for metric, label in WANT:
            if metric in col and label not in seen and r[col[metric]] not in ("", "n/a"):
                seen.add(label)
                print(f"  {label:38s} {r[col[metric]]} {units[col[metric]]}")


if __name__ == "__main__":
    main(sys.argv[1])
