"""Hot SASS lines (warp-stall samples) of one launch in an ncu report.
Usage: python tools/ncu_source_hot.py rep.ncu-rep <launch-skip> [min_pct]"""
import csv
import io
import subprocess
import sys


def main(path, skip, min_pct=1.0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    print(rows[0][:2])
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    print("total samples", tot, "instructions", len(data))
    for i, r in enumerate(data):
        n = int(r[isamp])
        if n > tot * min_pct / 100:
            st = sorted(((int(r[c]), hdr[c]) for c in stall_cols if r[c].isdigit()), reverse=True)[:2]
            print(f"{i:5d} {n:7d} {100 * n / tot:5.1f}%  {r[isrc][:90]:90s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
