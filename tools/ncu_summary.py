"""Key metrics of every kernel in an .ncu-rep as a markdown table (dev tool; summaries go to profiles/).

usage: ncu_summary.py <report.ncu-rep> [kernel-name-substring]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC / SM"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]
STALLS = "smsp__average_warps_issue_stalled_"  # ..._per_issue_active.ratio (newer ncu: smsp__average_warp_latency_issue_stalled_*)


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):  # some metrics carry a section prefix ("TPC.TriageCompute.<metric>")
        col.setdefault(h.split(".Triage")[-1].split(".", 1)[-1] if ".Triage" in h else h, i)
    ki = col["Kernel Name"]
    for n, r in enumerate(data):
        if sub not in r[ki]:
            continue
        print(f"### launch {n}: `{r[ki][:110]}`\n")
        print("| metric | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in col:
                print(f"| {label} (`{k}`) | {r[col[k]]} {units[col[k]]} |")
        stalls = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            tot = sum(v for v, _ in stalls) or 1.0
            print("| top stall reasons (warp-cycles per issue) | " +
                  ", ".join(f"{name} {v:.2f} ({100 * v / tot:.0f} %)" for v, name in stalls[:6]) + " |")
        print()


if __name__ == "__main__":
    main()
