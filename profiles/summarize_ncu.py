#!/usr/bin/env python
"""Turn an `ncu --set full` report (.ncu-rep, read here without a GPU) into the small JSON summary
committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/X.ncu-rep profiles/X_summary.json
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_red.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dm::", "")
        d = {"kernel": name, "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for m in METRICS:
            if m in idx:
                d[f"{m} [{units[idx[m]]}]"] = r[idx[m]]
        res.append(d)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    for d in res:
        print(d["kernel"], d.get("gpu__time_duration.sum [us]"))


if __name__ == "__main__":
    main()
