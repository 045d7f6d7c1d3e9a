#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -s KILL 1200 ncu --metrics $M --clock-control none -c 4000 --csv --log-file /tmp/all_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
python - <<'PY'
# keep the header + the last 2 complete steps (between image-packing launches) of the capture
import csv, io
txt = open("/tmp/all_cfg3.csv").read()
head = txt[:txt.index('"ID"')]
rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
hdr, body = rows[0], rows[1:]
ki, ii = hdr.index("Kernel Name"), hdr.index("ID")
ids = sorted({int(r[ii]) for r in body if len(r) > ki and "pack_image" in r[ki]})
print("launches captured:", max(int(r[ii]) for r in body if len(r) > ii) + 1, "pack_image at", ids)
lo = ids[-3] if len(ids) >= 3 else ids[0]
keep = [r for r in body if len(r) > ii and int(r[ii]) >= lo]
with open("gpurun_out/r2x_launches_cfg3.csv", "w", newline="") as f:
    f.write(head)
    w = csv.writer(f, quoting=csv.QUOTE_ALL)
    w.writerow(hdr)
    w.writerows(keep)
PY
python tools/step_table.py gpurun_out/r2x_launches_cfg3.csv | tail -70
