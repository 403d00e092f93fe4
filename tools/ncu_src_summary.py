"""Summarise an `ncu --page source --csv` export: opcode mix, hottest SASS lines, lane utilisation.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > src.csv
    python tools/ncu_src_summary.py src.csv [top_n]
"""
import csv
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ia, isrc, ith, ismp = (hdr.index(x) for x in ("Instructions Executed", "Source", "Avg. Threads Executed", "# Samples"))
    first = rows.index(hdr)
    later = [i for i in range(first + 1, len(rows)) if rows[i] == hdr]      # one table per launch: keep the first
    rows = rows[first:later[0]] if later else rows[first:]
    data = [r for r in rows if len(r) == len(hdr) and r[ia].isdigit()]
    tot = sum(int(r[ia]) for r in data)
    stot = sum(int(r[ismp]) for r in data)
    print("warp instructions", tot, "sass lines", len(data), "samples", stot)
    c, s = Counter(), Counter()
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[ia]); s[op] += int(r[ismp])
    for op, n in c.most_common(top):
        print(f"{op:10s} inst {n / tot * 100:5.1f}%  stall-samples {s[op] / max(stot, 1) * 100:5.1f}%")
    print("mean active threads per instruction", sum(int(r[ia]) * float(r[ith]) for r in data) / tot)
    print("--- hottest lines by samples")
    for r in sorted(data, key=lambda r: -int(r[ismp]))[:top]:
        print(f"{int(r[ismp]) / max(stot, 1) * 100:5.2f}%  exec {int(r[ia]):>12d} thr {float(r[ith]):4.1f}  {r[isrc].strip()[:90]}")


if __name__ == "__main__":
    main()
