"""Summarise .ncu-rep captures into profiles/<name>.md (run in the build container: `ncu -i` needs no GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "DRAM cycles active %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    return hdr, units, rd[2:]


def main(paths):
    for path in paths:
        hdr, units, rows = rows_of(path)
        print(f"## {path.split('/')[-1]}\n")
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            print(f"### `{name[:110]}`\n")
            print("| metric | value |\n|---|---|")
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("| top stall reasons (warps stalled per issue-active cycle) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " |")
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
