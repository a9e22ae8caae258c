"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --quick` into a per-kernel table
for ONE step of the hot path (crop .. assign_pnp).  Usage: python profiles/summarize_launches.py launches.csv"""
import collections
import csv
import re
import sys


def load(path, extra=None):
    """-> [(id, kernel name, us, grid)] in launch order; `extra` (dict) receives {id: {metric: value}} for any other
    metrics the capture holds (e.g. dram__bytes_read.sum, in bytes)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    seq, seen = [], {}
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        kid = int(row["ID"])
        name = row.get("Metric Name", "gpu__time_duration.sum")
        if name.startswith("gpu__time_duration"):
            v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
            seen[kid] = len(seq)
            seq.append((kid, row["Kernel Name"], v, row.get("Grid Size", "")))
        elif extra is not None:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            extra.setdefault(kid, {})[name] = v * scale
    return seq


def short(name):
    n = re.sub(r"\(.*", "", name)
    n = re.sub(r"^.*::", "", n)
    for fam in ("gemm_tc2_kernel", "gemm_tc_kernel", "conv3_tc_kernel", "ffn_tc2_kernel", "ffn_tc_kernel"):
        if fam in name:
            return fam
    return "gemm_tc_kernel" if n.strip().startswith("GemmKParams") else n.strip()


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
