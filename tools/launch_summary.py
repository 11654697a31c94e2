"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (dev tool).

    python tools/launch_summary.py gpurun_out/launches.csv [first_id last_id]  > profiles/rNN_launch_list_step.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iid, iname, imetric, ival, iunit = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    for r in rd:
        if r[imetric] != "gpu__time_duration.sum":
            continue
        k = int(r[iid])
        if lo <= k <= hi:
            v = float(r[ival].replace(",", ""))
            unit = r[iunit]
            v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
            rows.append((k, r[iname], v_ms))
    agg = defaultdict(lambda: [0.0, 0])
    for _, name, ms in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"<.*", lambda m: m.group(0)[:24], short)
        agg[short][0] += ms
        agg[short][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"launches {len(rows)}, sum of kernel durations {total:.2f} ms (cold-cache, serialised)")
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:9.3f} ms {n:5d} {100 * ms / total:5.1f}%  {name[:90]}")


if __name__ == "__main__":
    main()
