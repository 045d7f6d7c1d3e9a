"""Print one complete step (from the last-but-one image-pack launch to the next) of an ncu launch list
(`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`): per launch time, DRAM MB,
GB/s; then per-kernel-name totals.   python tools/step_table.py profiles/r2k_launches_cfg3.csv [--totals-only]"""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    ki, mi, vi, ii, gi = (h.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Grid Size'))
    d = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        k = int(r[ii])
        d.setdefault(k, {'name': r[ki], 'grid': r[gi]})
        d[k][r[mi]] = float(r[vi].replace(',', ''))
    return d


def short(name):
    name = re.sub(r'\(.*', '', name)
    return name.replace('void ', '').replace('unnamed>::', '').replace('ssdsb::<', '').replace('ssdsb::', '')[:44]


def main():
    d = load(sys.argv[1])
    ids = sorted(d)
    starts = [i for i in ids if 'pack_image' in d[i]['name']]
    s, e = starts[-2], starts[-1]
    tot, totals = 0.0, OrderedDict()
    for i in ids:
        if i < s or i >= e:
            continue
        k = d[i]
        t = k.get('gpu__time_duration.sum', 0) / 1000
        b = (k.get('dram__bytes_read.sum', 0) + k.get('dram__bytes_write.sum', 0)) / 1e6
        tot += t
        nm = short(k['name'])
        a = totals.setdefault(nm, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += b
        if '--totals-only' not in sys.argv:
            print(f"{i - s:3d} {nm:44s} {k['grid']:>14s} {t:7.1f} us {b:7.1f} MB {b / t if t else 0:5.2f} TB/s  cum {tot:7.1f}")
    print(f"\n{e - s} launches, {tot:.1f} us (cold, serialised)")
    for nm, (n, t, b) in sorted(totals.items(), key=lambda kv: -kv[1][1]):
        print(f"  {nm:44s} x{n:<3d} {t:8.1f} us {b:8.1f} MB")


if __name__ == "__main__":
    main()
