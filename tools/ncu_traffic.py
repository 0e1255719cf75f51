#!/usr/bin/env python
"""DRAM traffic per analysed row and stage from an `ncu --set full` raw CSV of tools/profile_step.py.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv 512 > profiles/r01_traffic.json

bench.py reads the file to fill roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of the dominant
kernel, scaled to the rows one bench launch processes).
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def stage_of(name: str):
    if "edc_kernel" in name:
        return "edc"
    if "fft_rows_untangle_reg" in name:
        return "fft_rows_untangle"
    if "untangle_kernel" in name:
        return "untangle"
    if "band_fit_kernel" in name:
        return "band_fit"
    if "band_fused_kernel" in name:
        return "band_fused"
    if "csd_tma_kernel" in name:
        return "csd"
    m = re.search(r"fft_cols_reg<.*>, (\d+), \d+>", name)
    if m:
        return "fft_cols_band" if m.group(1) == "3" else "fft_cols"
    m = re.search(r"fft_rows_reg<.*>, \d+, (\d)>", name)
    if m:
        return "fft_rows_band" if m.group(1) == "1" else "fft_rows"
    return None


rows = list(csv.reader(open(sys.argv[1])))
n_rows = int(sys.argv[2])
hdr, units = rows[0], rows[1]
kn, rd, wr, du = (hdr.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
out = {}
for r in rows[2:]:
    st = stage_of(r[kn])
    if st is None:
        continue
    b = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
    e = out.setdefault(st, {"dram_bytes_per_row": 0.0, "kernels": 0, "ncu_us_per_row": 0.0})
    e["dram_bytes_per_row"] += b / n_rows
    e["ncu_us_per_row"] += float(r[du]) / n_rows
    e["kernels"] += 1
for e in out.values():
    e["dram_bytes_per_row"] = round(e["dram_bytes_per_row"], 1)
    e["ncu_us_per_row"] = round(e["ncu_us_per_row"], 4)
print(json.dumps({"source": "ncu --set full --clock-control none, tools/profile_step.py --rows %d --steps 1 (one launch of every "
                            "kernel of the step; dram__bytes_read.sum + dram__bytes_write.sum)" % n_rows,
                  "rows": n_rows, "stages": out}, indent=1))
