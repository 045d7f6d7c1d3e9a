"""Group an ncu launch list (bench.py --steps 1) by layer group of the cfg-2 model.
    python tools/launch_groups.py gpurun_out/launches_A.csv [gpurun_out/launches_B.csv]"""
import csv
import sys


def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    rows = []
    for row in r:
        if len(row) < len(hdr):
            continue
        rows.append((row[ix['Kernel Name']], float(row[ix['Metric Value']].replace(',', ''))))
    idx = [i for i, r in enumerate(rows) if 'pack_image' in r[0]]
    return rows[idx[-2]:idx[-1]]


names = ['pack', 'stem', 'maxpool'] + ['l1.down', 'l1.0.c1', 'l1.0.c2', 'l1.0.c3'] + \
    [f'l1.{i}.c{j}' for i in (1, 2) for j in (1, 2, 3)]
names += ['l2.down', 'l2.0.c1', 'l2.0.c2', 'l2.0.c3'] + [f'l2.{i}.c{j}' for i in (1, 2, 3) for j in (1, 2, 3)]
names += ['l3.down', 'l3.0.c1', 'l3.0.c2', 'l3.0.c3'] + [f'l3.{i}.c{j}' for i in range(1, 6) for j in (1, 2, 3)]
names += ['l4.down', 'l4.0.c1', 'l4.0.c2', 'l4.0.c3'] + [f'l4.{i}.c{j}' for i in (1, 2) for j in (1, 2, 3)]
names += ['ex0.a', 'ex0.b', 'ex1.a', 'ex1.b', 'ex2.a', 'ex2.b'] + [f'head{i}' for i in range(6)]
names += ['dec_select', 'dec_final', 'nms', 'torch1', 'torch2', 'torch3']


def groups(rows):
    g = {}
    for i, (k, v) in enumerate(rows):
        nm = names[i] if i < len(names) else '?'
        key = nm[:2] + '.' + nm.split('.')[-1] if nm.startswith('l') and '.' in nm else nm.split('.')[0]
        g[key] = g.get(key, 0) + v / 1e3
    return g


gs = [groups(load(f)) for f in sys.argv[1:]]
for k in gs[0]:
    print(f"{k:12s} " + "  ".join(f"{g.get(k, 0):9.1f}" for g in gs))
print(f"{'total':12s} " + "  ".join(f"{sum(g.values()):9.1f}" for g in gs))
