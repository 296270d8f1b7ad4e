#!/usr/bin/env python
"""From an `ncu --set full --page raw --csv` export of ONE extractor step (tools/prof_extract.py B steps), write per-stage
per-FRAME DRAM bytes (profiles/traffic.json) and thread instructions (profiles/issue.json), which bench.py reads for
`roofline.traffic` and `roofline_issue`.   usage: ncu_stage_json.py raw.csv FRAMES source-note"""
import csv
import json
import os
import sys

STAGE = [("k_pyr_level", "pyramid"), ("k_fast_cells", "fast_cells"), ("k_qt_", "quadtree"), ("k_quadtree", "quadtree"), ("k_assemble", "assemble"),
         ("k_blur", "blur"), ("k_orient_desc", "orient_desc")]
rows = list(csv.reader(open(sys.argv[1])))
frames = int(sys.argv[2])
note = sys.argv[3] if len(sys.argv) > 3 else ""
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
units = dict(zip(hdr, rows[1]))


def val(r, k):
    v = float(r[ix[k]].replace(",", ""))
    u = units.get(k, "")
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)


traffic, issue, times = {}, {}, {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    st = next((s for k, s in STAGE if k in name), None)
    if st is None:
        continue
    traffic[st] = traffic.get(st, 0) + (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / frames
    issue[st] = issue.get(st, 0) + val(r, "smsp__inst_executed.sum") * val(r, "smsp__thread_inst_executed_per_inst_executed.ratio") / frames
    times[st] = times.get(st, 0) + val(r, "gpu__time_duration.sum")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
traffic = {k: int(v) for k, v in traffic.items()}
issue = {k: int(v) for k, v in issue.items()}
traffic["_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum per FRAME per stage ({note}); writes mostly stay in the 126 MB L2 during a launch"
issue["source"] = f"smsp__inst_executed.sum x smsp__thread_inst_executed_per_inst_executed.ratio per FRAME per stage ({note})"
json.dump(traffic, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
json.dump(issue, open(os.path.join(root, "profiles", "issue.json"), "w"), indent=1)
print("per-frame DRAM bytes:", traffic)
print("per-frame thread instructions:", issue)
print("ncu time share:", {k: round(v / sum(times.values()), 3) for k, v in times.items()})
