"""ncu launch list with DRAM byte counters -> profiles/r2_traffic.json, the record bench.py's `roofline.traffic` reads.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/X.csv python tools/prof_run.py --windows 1010 --max-batch 1005
    python tools/traffic_record.py gpurun_out/X.csv 1010 f16x3 > profiles/r2_traffic.json

The record carries the sha1 of the CUDA sources it was measured on (bench.kernel_source_hash); bench.py reports
`traffic: null` when that hash is not the hash of the sources it runs, so a kernel change cannot leave a stale number.
"""
import collections
import csv
import datetime
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

FAMILY = {"features_kernel": "features", "conv_tc_kernel": "classifier", "conv1_direct": "classifier",
          "pool_planar": "classifier", "mask_head_partials": "classifier", "mask_head_planar": "classifier",
          "average_kernel": "postproc", "regions_kernel": "postproc", "scan_counts_kernel": "postproc"}


def main(path, windows, mode):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per_id = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per_id.setdefault(row["ID"], {"name": row["Kernel Name"]})
        try:
            d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    fam = collections.defaultdict(lambda: {"launches": 0, "bytes": 0.0, "us": 0.0})
    for d in per_id.values():
        key = next((v for k, v in FAMILY.items() if k in d["name"]), "other")
        fam[key]["launches"] += 1
        fam[key]["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        fam[key]["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    rec = {"kernel_source_sha1": bench.kernel_source_hash(), "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ"),
           "how": f"one pass of {windows} windows (tools/prof_run.py --windows {windows} --max-batch 1005), all launches of a "
                  "kernel family summed, / windows", "windows": windows,
           "dram_bytes_per_window": {mode: {k: v["bytes"] / windows for k, v in fam.items()}},
           "launches": {k: v["launches"] for k, v in fam.items()},
           "ncu_time_share": {k: round(v["us"] / max(sum(x["us"] for x in fam.values()), 1e-9), 4) for k, v in fam.items()}}
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "f16x3")
