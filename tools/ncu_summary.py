"""Summarise an Nsight Compute report (read here, on the CPU box) into a small markdown table for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep [more.ncu-rep ...] > profiles/ncu_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    return header, units, rows[2:]


def main():
    for path in [a for a in sys.argv[1:] if not a.startswith('--')]:
        header, units, rows = load(path)
        col = {h: i for i, h in enumerate(header)}
        name_i = col.get("Kernel Name")
        print(f"## {path}\n")
        seen = set()
        for r in rows:
            if r[name_i] in seen and "--all" not in sys.argv:      # one capture per distinct kernel is enough
                continue
            seen.add(r[name_i])
            print(f"### `{r[name_i][:110]}`\n")
            print("| metric | value |\n|---|---|")
            for key, label in KEYS:
                hit = [h for h in header if h == key or h.startswith(key)]
                if hit:
                    i = col[hit[0]]
                    print(f"| {label} (`{hit[0]}`) | {r[i]} {units[i]} |")
            stalls = []
            for h in header:
                if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[col[h]].replace(",", "")), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                print("| top stall reasons (warps stalled per issue-active cycle) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " |")
            print()


if __name__ == "__main__":
    main()
