"""Summarise an .ncu-rep (from `ncu --set full`) into a small markdown table: python tools/ncu_summary.py rep.ncu-rep > out.md"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thr"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__cycles_active.avg", "cycles")]


def main():
    rep = sys.argv[1]
    # a .ncu-rep, or the CSV `ncu -i rep --page raw --csv` exported on the GPU box (the reports exceed gpurun's copy-back limit)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{rep.split('/')[-1]}` (per launch; cold-cache, serialised)\n")
    cols = [(m, n) for m, n in WANT if m in idx]
    print("| kernel | " + " | ".join(f"{n} [{units[idx[m]]}]" if units[idx[m]] else n for m, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for m, _ in cols:
            v = r[idx[m]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
