"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --quick` into a per-kernel table
for ONE step of the hot path (crop .. assign_pnp).  Usage: python profiles/summarize_launches.py launches.csv"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        seq.append((int(row["ID"]), row["Kernel Name"], v, row.get("Grid Size", "")))
    return seq


def short(name):
    n = re.sub(r"\(.*", "", name)
    n = re.sub(r"^.*::", "", n)
    return "gemm_tc_kernel" if n.strip().startswith("GemmKParams") or "gemm_tc" in name else n.strip()


def main(path):
    seq = load(path)
    starts = [i for i, s in enumerate(seq) if "crop_resize" in s[1]]
    ends = [i for i, s in enumerate(seq) if "assign_pnp" in s[1] or "PnpDesc" in s[1]]
    a = starts[-1] if ends and ends[-1] > starts[-1] else starts[-2]
    b = min(e for e in ends if e > a)
    step = seq[a:b + 1]
    tot = sum(s[2] for s in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s in step:
        k = short(s[1])
        agg[k][0] += 1
        agg[k][1] += s[2]
    print(f"one step = {len(step)} launches, {tot / 1000:.3f} ms of kernel time (ncu: serialised, cold caches)\n")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
    print("\ntop 12 launches:\n")
    print("| id | kernel | us | grid |\n|---|---|---:|---|")
    for s in sorted(step, key=lambda s: -s[2])[:12]:
        print(f"| {s[0]} | `{short(s[1])}` | {s[2]:.1f} | {s[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
