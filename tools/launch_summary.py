"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the run."""
import collections
import csv
import re
import sys


def main(path, iters):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("rg::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total, {iters} iterations "
          f"(cold-cache serialised ncu times: compare SHARES, not absolutes)")
    print(f"{'share':>7} {'ms/iter':>9} {'n/iter':>7} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[1] / iters / 1000:9.3f} {v[0] / iters:7.1f} {v[1] / v[0]:9.1f}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
