"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv` launch list by kernel (template arguments kept)."""
import collections
import csv
import re
import sys


def short(name):
    base = name.split("(")[0]
    m = re.search(r"(conv_tc_kernel<[^>]*>)", name)
    if m:
        return m.group(1)
    return base.split("::")[-1].split("<")[0][:40]


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_id = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per_id.setdefault(row["ID"], {"name": short(row["Kernel Name"])})
        try:
            d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for d in per_id.values():
        a = agg.setdefault(d["name"], {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "tp": 0.0})
        a["n"] += 1
        t = d.get("gpu__time_duration.sum", 0.0) / 1e3
        a["t"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
    tot = sum(a["t"] for a in agg.values())
    print(f"{len(per_id)} launches, {tot:.1f} us total (cold-cache serialised launch times: compare shares)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        print(f"{k:36s} n={a['n']:3d} total={a['t']:9.1f} us share={a['t'] / tot * 100:5.1f}%  dram rd={a['rd'] / 1e6:9.1f} MB "
              f"wr={a['wr'] / 1e6:9.1f} MB  ({(a['rd'] + a['wr']) / max(a['t'], 1e-9) / 1e6:6.2f} TB/s)  tensor-pipe active "
              f"{a['tp'] / max(a['t'], 1e-9):5.1f}%")
    print(f"DRAM total: {sum(a['rd'] + a['wr'] for a in agg.values()) / 1e9:.2f} GB")


if __name__ == "__main__":
    main(sys.argv[1])
