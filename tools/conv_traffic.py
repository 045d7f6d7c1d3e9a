"""DRAM traffic of the conv launches of ONE step from an ncu launch list of `bench.py --config X --steps 2 ...`
(metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum): writes the JSON bench.py reads for
`roofline.traffic`.   python tools/conv_traffic.py <launches.csv> <config name> <out.json>"""
import collections
import csv
import io
import json
import sys

path, name, out = sys.argv[1:4]
txt = open(path).read()
rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
hdr = rows[0]
by = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    e = by.setdefault(d["ID"], {"name": d["Kernel Name"]})
    e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
items = list(by.values())
idx = [i for i, d in enumerate(items) if "pack_image" in d["name"]]
step = items[idx[-2]:idx[-1]]                      # one complete step between two image-packing launches
conv = [d for d in step if any(k in d["name"] for k in ("conv_igemm", "conv_pair", "dwconv", "mbconv", "maxpool", "upsample",
                                                         "bifpn_fuse"))]
tot_t = sum(d["gpu__time_duration.sum"] for d in step) / 1e3
res = {"config": name,
       "source": f"profiles/{path.split('/')[-1]} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                 "dram__bytes_write.sum --clock-control none, bench.py --config %s --steps 2 --warmup 3)" % name,
       "conv_launches_per_step": len(conv),
       "conv_dram_bytes_per_step": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in conv),
       "conv_time_us_cold_serialised": sum(d["gpu__time_duration.sum"] for d in conv) / 1e3,
       "step_time_us_cold_serialised": tot_t}
res["conv_share_of_step"] = res["conv_time_us_cold_serialised"] / tot_t
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
