#!/usr/bin/env python
"""Summarise an ncu capture of ONE chunk step (bench.py --profile-step under `ncu --profile-from-start off`).

    python tools/ncu_step_summary.py gpurun_out/step.ncu-rep gpurun_out/kernel_labels.json profiles/r01_step_<tag>

Writes <out>.csv (one row per launch: label, SASS kernel, duration, DRAM bytes, tensor-pipe %, L2 %) and refreshes
profiles/ncu_traffic.json (DRAM bytes per launch by kernel class) which bench.py reports as roofline.traffic.
Runs here (no GPU): it only reads the report with `ncu -i`.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, labels_path, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
meta = json.load(open(labels_path))
labels = meta["labels"]
data = rows[2:]
assert len(data) == len(labels), f"{len(data)} launches captured, {len(labels)} labels"
cols = {
    "duration_us": ("gpu__time_duration.sum", 1e-3),
    "dram_read_bytes": ("dram__bytes_read.sum", None),
    "dram_write_bytes": ("dram__bytes_write.sum", None),
    "tensor_pipe_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "l2_sectors_pct": ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", 1),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
}
units = rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}


def val(r, metric):
    i = idx.get(metric)
    if i is None or r[i] == "":
        return None
    return float(r[i].replace(",", "")) * scale.get(units[i], 1)


table, traffic, count = [], {}, {}
for lab, r in zip(labels, data):
    rec = {"label": lab, "kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")}
    for k, (m, mult) in cols.items():
        v = val(r, m)
        rec[k] = None if v is None else (v * 1e-3 if k == "duration_us" else v)
    table.append(rec)
    key = re.sub(r"step\d+", "step", lab)
    if rec["dram_read_bytes"] is not None:
        traffic[key] = traffic.get(key, 0.0) + rec["dram_read_bytes"] + rec["dram_write_bytes"]
        count[key] = count.get(key, 0) + 1
with open(out + ".csv", "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=list(table[0].keys()))
    w.writeheader()
    w.writerows(table)
tot = sum(t["duration_us"] for t in table)
print(f"{len(table)} launches, {tot:.1f} us under ncu (cold-cache, serialised: compare shares)")
for t in sorted(table, key=lambda t: -t["duration_us"])[:12]:
    print(f"  {t['label']:34s} {t['duration_us']:8.1f} us  {100 * t['duration_us'] / tot:5.1f}%  dram "
          f"{(t['dram_read_bytes'] or 0) / 1e6:7.1f}+{(t['dram_write_bytes'] or 0) / 1e6:6.1f} MB  tensor {t['tensor_pipe_pct']}")
tj = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_enhancement_mi_b200 import build as native_build  # noqa: E402
# the capture is only valid for the library it was taken on: bench.py reports `traffic` when this hash matches its own
json.dump({"source": os.path.basename(out) + ".csv", "capture": os.path.basename(out), "source_hash": native_build.source_hash(),
           "streams": meta["streams"], "precision": meta["precision"],
           "dram_bytes_per_launch": {k: traffic[k] / count[k] for k in traffic}}, open(tj, "w"), indent=1)
