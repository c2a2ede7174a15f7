"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel + grid."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        m = re.search(r"(conv_tc_kernel<[^>]*>|conv_nhwc_f32<[\d, ]+>|[A-Za-z_0-9]+)(?=\(|<|$)", name.split("::")[-1])
        name = m.group(1) if m else name[:50]
        t = float(row["Metric Value"]) / 1e3
        key = (name, row.get("Grid Size", ""), row.get("Block Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"{n} launches, {tot:.1f} us total")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[0]:28s} grid={k[1]:16s} n={a[0]:3d} total={a[1]:9.1f} us avg={a[1] / a[0]:8.1f} us share={a[1] / tot * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
