"""Blackwell-native evidence, from the built library (no GPU needed):

    python tools/sass_summary.py > profiles/r2_sass_summary.txt

Per kernel of softspoken_b200/libsoftspoken_b200.so: counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA
path (`cuobjdump -sass`: UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk,
UTMALDG = tensor-map TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync, FFMA2 / FADD2 / FMUL2 = the packed fp32 pair
instructions of sm_100) and the registers,
spills and static shared memory ptxas reported for it (softspoken_b200/csrc/*.ptxas.log).
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "softspoken_b200", "libsoftspoken_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "REDUX", "SHFL",
             "FFMA2", "FADD2", "FMUL2"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(MNEMONICS, 0)
            counts[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["instructions"] += 1
            if op in counts[cur]:
                counts[cur][op] += 1
    res = {}
    for log in glob.glob(os.path.join(ROOT, "softspoken_b200", "csrc", "*.ptxas.log")):
        fn, props_of = None, None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                fn = m.group(1)
                res[fn] = {}
            m = re.search(r"Function properties for (\S+)", line)
            if m:
                props_of = m.group(1)
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and fn and props_of == fn:
                res[fn]["spill"] = f"{m.group(1)}/{m.group(2)}"
            m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
            if m and fn:
                res[fn]["regs"] = m.group(1)
                res[fn]["smem"] = m.group(2) or "0"
    names = demangle(list(counts))
    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# SASS / ptxas summary of libsoftspoken_b200.so (sm_100a), tree at {git}+; columns: instructions, "
          + ", ".join(MNEMONICS) + ", registers, spill st/ld bytes, static smem")
    tot = dict.fromkeys(MNEMONICS, 0)
    for fn in sorted(counts, key=lambda f: names[f]):
        c, r = counts[fn], res.get(fn, {})
        for k in MNEMONICS:
            tot[k] += c[k]
        short = names[fn].replace("(anonymous namespace)::", "").replace("void ", "").replace("ss::tc::Prec", "")
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)      # drop the parameter list
        print(f"{short[:86]:86s} {c['instructions']:6d} " + " ".join(f"{c[k]:5d}" for k in MNEMONICS)
              + f"  {r.get('regs', '?'):>4s} {r.get('spill', '?'):>9s} {r.get('smem', '?'):>6s}")
    print("TOTAL".ljust(86) + "        " + " ".join(f"{tot[k]:5d}" for k in MNEMONICS))
    assert tot["UTCHMMA"] > 0 and tot["LDTM"] > 0 and tot["UBLKCP"] > 0 and tot["HMMA"] == 0, tot


if __name__ == "__main__":
    sys.exit(main())
