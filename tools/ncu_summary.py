"""Summarise an `ncu --set full` report (.ncu-rep) per launch: duration, DRAM traffic, tensor-pipe activity, L2.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.json]"""
import csv
import json
import subprocess
import sys

FIELDS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_bytes"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_elapsed.max", "cycles"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", ""), "grid": r[idx["Grid Size"]]}
        for key, name in FIELDS:
            if key in idx:
                v = float(r[idx[key]].replace(",", "") or 0)
                d[name] = v * SCALE.get(units[idx[key]], 1.0)
        launches.append(d)
    print(f"# {path}: {len(launches)} profiled launches (ncu --set full --clock-control none)")
    print(f"{'kernel':34s} {'grid':>12s} {'us':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'tensor%':>8s} {'L2->SM MB':>10s} "
          f"{'L2 hit%':>8s} {'L2 thr%':>8s}")
    for d in launches:
        print(f"{d['kernel'][:34]:34s} {d['grid']:>12s} {d.get('duration', 0):8.1f} {d.get('dram_read', 0) / 1e6:9.1f} "
              f"{d.get('dram_write', 0) / 1e6:9.1f} {d.get('tensor_active_pct', 0):8.1f} "
              f"{d.get('l2_to_sm_bytes', 0) / 1e6:10.1f} {d.get('l2_hit_pct', 0):8.1f} {d.get('l2_throughput_pct', 0):8.1f}")
    if out:
        agg = {}
        for d in launches:
            a = agg.setdefault(d["kernel"], {"launches": 0, "dram_bytes": 0.0, "duration_us": 0.0, "tensor_active_pct": 0.0})
            a["launches"] += 1
            a["dram_bytes"] += d.get("dram_read", 0) + d.get("dram_write", 0)
            a["duration_us"] += d.get("duration", 0)
            a["tensor_active_pct"] += d.get("tensor_active_pct", 0)
        for a in agg.values():
            n = a["launches"]
            a["dram_bytes_per_launch"] = a.pop("dram_bytes") / n
            a["duration_us_per_launch"] = a.pop("duration_us") / n
            a["tensor_active_pct"] /= n
        json.dump({"source": path, "kernels": agg}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
