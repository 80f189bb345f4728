"""Attribute the per-instruction columns of an ncu source page (SASS view, CSV) to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g <cubin> > k.sass
    python tools/ncu_by_line.py src.csv k.sass <mangled kernel name> [file filter]

The ncu CSV of this toolkit carries metrics only in the SASS view; nvdisasm -g prints '//## File "..", line N' markers
in front of the instructions of the same function in the same order, so the i-th instruction of one is the i-th of the other.
"""
import collections
import csv
import re
import sys


def sass_lines(path, kernel):
    out, on, cur = [], False, None
    for ln in open(path):
        if ln.startswith(".text."):
            on = ln.strip() == f".text.{kernel}:"
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    filt = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = sass_lines(sass, kernel)
    inst = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["Address"]].startswith("0x")]
    n = min(len(lines), len(inst))
    print(f"# sass instructions: nvdisasm {len(lines)}, ncu {len(inst)}", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0])
    ti = ts = 0
    for k in range(n):
        key = lines[k]
        if key is None:
            key = ("?", 0)
        a = agg[key]
        i_ = int(inst[k][ix["Instructions Executed"]] or 0)
        s_ = int(inst[k][ix["# Samples"]] or 0)
        a[0] += i_
        a[1] += s_
        ti += i_
        ts += s_
    print(f"total instructions {ti}, samples {ts}")
    for (f, l), (i_, s_) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if filt and filt not in f:
            continue
        if i_ * 200 < ti and s_ * 200 < ts:
            continue
        print(f"{f}:{l:5d}  inst {100 * i_ / ti:5.2f}%  samples {100 * s_ / ts:5.2f}%")


if __name__ == "__main__":
    main()
