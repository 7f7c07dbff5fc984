#!/usr/bin/env python
"""profiles/traffic.json: DRAM bytes per launch of every bench stage, from the committed `ncu --set full` summaries
(profiles/summarize_ncu.py output) of ONE step of `python bench.py` at the default batch (256 frames).

    python profiles/make_traffic.py 256 profiles/r4/ncu_full_step.csv
"""
import csv
import json
import os
import sys

STAGE_OF = [("leaf_band", "leaf_stats"), ("leaf_median", "median"), ("edt_argmax", "edt_rows"), ("select_leaf", "select"),
            ("chamfer", "chamfer"), ("outside_max", "orientation"), ("leaf_boundary", "orientation"), ("orient", "orientation"),
            ("score_kernel", "score_maps"), ("nms_tiles", "candidates"), ("nms_kernel", "candidates"), ("gather_kernel", "patches"), ("compact_slots", "patches"),
            ("conv3x3_umma", "cnn"), ("pool2x2", "cnn"), ("pack_input", "cnn"), ("cnn_tail", "cnn"), ("fuse_kernel", "fuse")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(frames, files):
    out = {}
    seen = set()
    for f in files:
        rows = list(csv.reader(open(f)))
        hdr = rows[0]
        col = {h.split(" [")[0]: (i, h.split("[")[-1].rstrip("]") if "[" in h else "") for i, h in enumerate(hdr)}
        for r in rows[1:]:
            name = r[col["Kernel Name"][0]]
            stage = next((s for k, s in STAGE_OF if k in name), None)
            if stage is None:
                continue
            key = (name, r[col["launch__grid_size"][0]])
            if stage != "cnn" and key in seen:      # the capture may hold a kernel of two consecutive steps
                continue
            seen.add(key)
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i, unit = col[m]
                tot += float(r[i]) * UNIT.get(unit, 1.0)
            e = out.setdefault(stage, {"frames": frames, "dram_bytes_per_launch": 0.0, "kernels": []})
            e["dram_bytes_per_launch"] += tot
            e["kernels"].append(name.split("(")[0].replace("<unnamed>::", "").replace("void ", ""))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    for k, v in out.items():
        print(f"{k:12s} {v['dram_bytes_per_launch'] / 1e6:10.1f} MB  {len(v['kernels'])} launches")


if __name__ == "__main__":
    main(int(sys.argv[1]), sys.argv[2:])
