"""profiles/rNN_sass_summary.txt: per kernel of the built library, the counts of the SASS mnemonics that identify the
Blackwell-native paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk),
UTMALDG (tensor-map TMA), LDGSTS (cp.async), HMMA (mma.sync), packed fp32x2 arithmetic, 256-bit global loads.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uforecon_b200", "libuforecon_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "HMMA", "FFMA2", "FADD2", "FMUL2", "LDG.E.ENL2.256", "SHFL", "MUFU", "BAR.SYNC", "SYNCS"]


def main():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, capture_output=True)
        rows = []
        for f in sorted(os.listdir(d)):
            if not f.endswith(".cubin") or "sm_100a" not in f:
                continue
            sass = subprocess.run(["cuobjdump", "-sass", os.path.join(d, f)], capture_output=True, text=True).stdout
            cur, cnt, n = None, None, 0
            for ln in sass.splitlines():
                m = re.search(r"Function : (\S+)", ln)
                if m:
                    if cur:
                        rows.append((cur, n, cnt))
                    cur, cnt, n = m.group(1), collections.Counter(), 0
                    continue
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
                if m and cur:
                    n += 1
                    op = m.group(1)
                    for k in KEYS:
                        if op.startswith(k):
                            cnt[k] += 1
            if cur:
                rows.append((cur, n, cnt))
    dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel of uforecon_b200/libuforecon_b200.so (cuobjdump -sass, sm_100a)")
    print("# kernel | instructions | " + " | ".join(KEYS))
    for (name, n, cnt), dn in sorted(zip(rows, dem), key=lambda x: x[1]):
        short = (dn[:dn.find(">(") + 1] if ">(" in dn else dn.split("(")[0]).replace("void ufo::", "").replace("ufo::", "").replace("(int)", "").replace("(bool)", "")
        if not any(cnt.values()) and n < 400:
            continue
        print(f"{short} | {n} | " + " | ".join(str(cnt.get(k, 0)) for k in KEYS))


if __name__ == "__main__":
    main()
