#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel:
    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_bench_launches.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "usecond": v, "nsecond": v / 1e3, "ms": v * 1e3, "msecond": v * 1e3}.get(r[iu], v)
    tot[r[ik][:78]] += v
    cnt[r[ik][:78]] += 1
total = sum(tot.values())
print(f"{'kernel':<80}{'launches':>9}{'total us':>11}{'avg us':>9}{'share':>7}")
for k, v in tot.most_common():
    print(f"{k:<80}{cnt[k]:>9}{v:>11.1f}{v / cnt[k]:>9.1f}{100 * v / total:>6.1f}%")
print(f"{'total':<80}{sum(cnt.values()):>9}{total:>11.1f}")
