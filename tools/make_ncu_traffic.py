#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of one force iteration (tools/gpu/profile.sh):
dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the names bench.py gives the kernels.

    python tools/make_ncu_traffic.py gpurun_out/r2_full.ncu-rep "ball_h0=0.02" profiles/ncu_traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

NAMES = {"prep_kernel": "prep(zero+pad)", "cull_scatter_kernel": "cull_scatter", "adjacency_kernel": "adjacency",
         "cull_bin_kernel": "cull_bin", "tile_rows_kernel": "tile_rows", "vertex_update_kernel": "vertex_update+maxdp", "project_list_kernel": "project_escaped"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, workload, out = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("<")[0].split("(")[0].replace("void ", "").replace("dm::", "").strip()
        key = NAMES.get(name)
        if key is None or key in res:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]]) * UNIT[units[idx[m]]]
        res[key] = int(tot)
    data = {}
    if os.path.exists(out):
        with open(out) as f:
            data = json.load(f)
    data[workload] = res
    data.setdefault("_source", {})[workload] = os.path.basename(rep)
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
