#!/usr/bin/env python
"""Writes profiles/traffic.json from an `ncu --set full` report of the fused pool kernel: DRAM bytes (read + write) of one
launch and the frames that launch scored, keyed by rig.  bench.py scales it per frame into roofline.traffic.
    python tools/ncu_traffic.py gpurun_out/X.ncu-rep <views> <joints> <frames_per_launch>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path, views, joints, frames):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "score_pool_fused_kernel" not in d.get("Kernel Name", ""):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[k].replace(",", "")) * UNIT[units[hdr.index(k)]]
        recs.append((tot, float(d["gpu__time_duration.sum"].replace(",", "")), d["Kernel Name"].split("(")[0]))
    assert recs, "no fused kernel launch in the report"
    tot, ms, name = recs[-1]
    dst = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(dst)) if os.path.isfile(dst) else {}
    table["score_pool_fused_kernel_v%d_j%d" % (views, joints)] = {
        "dram_bytes": tot, "frames": frames, "kernel": name, "launch_ms_under_ncu": ms,
        "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of one launch (%s)"
                  % os.path.basename(path)}
    json.dump(table, open(dst, "w"), indent=1)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
