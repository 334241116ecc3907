#!/usr/bin/env python3
"""Static resource report of every kernel in librnagan_b200.so, without a GPU: registers, stack frame, spills and static
shared memory from `nvcc -Xptxas -v`, plus how often the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA,
TMA loads / stores, TMEM loads, programmatic dependent launch) occur per kernel (`cuobjdump -sass`).

    python tools/ptxas_report.py > profiles/r2_ptxas_report.txt
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rnagan_b200", "csrc")
SOURCES = ["rg_gemm_api.cu", "rg_ops.cu", "rg_img.cu", "rg_data.cu"]
MNEMONICS = {"UTCMMA": "UTC(H|Q|O)?MMA", "UTMALDG": "UTMALDG", "UTMASTG": "UTMASTG", "LDTM": "LDTM", "ACQBULK": "ACQBULK",
             "PREEXIT": "PREEXIT", "HMMA": r"HMMA\.", "LDSM": "LDSM", "LDGSTS": "LDGSTS"}


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o).replace("void ", "").replace("rg::", "") for o in out[:len(names)]]


def main():
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "gpurun_out") if os.path.isdir(
            os.path.join(ROOT, "gpurun_out")) else None) as tmp:
        so = os.path.join(tmp, "lib.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v",
               "-shared", "-Xcompiler", "-fPIC", "-o", so] + [os.path.join(CSRC, s) for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stderr)
            raise SystemExit("nvcc failed")
        sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    rows, cur = {}, None
    for ln in res.stderr.split("\n"):
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            cur = m.group(1)
            rows[cur] = {"regs": 0, "stack": 0, "spill_st": 0, "spill_ld": 0, "smem": 0}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m:
            rows[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", ln)
        if m:
            rows[cur]["regs"] = int(m.group(1))
            m2 = re.search(r"(\d+) bytes smem", ln)
            rows[cur]["smem"] = int(m2.group(1)) if m2 else 0
    counts, cur = {}, None
    for ln in sass.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = {k: 0 for k in MNEMONICS}
            continue
        if cur:
            for k, pat in MNEMONICS.items():
                if re.search(r"\b" + pat + r"\b", ln):
                    counts[cur][k] += 1
    names = sorted(rows)
    pretty = demangle(names)
    print(f"# {len(names)} kernels, sm_100a, nvcc -O3; spills > 0 are flagged with '!'")
    print(f"{'regs':>5} {'stack':>6} {'spillB':>7} {'smemB':>6}  " + " ".join(f"{k:>8}" for k in MNEMONICS) + "  kernel")
    for n, p in sorted(zip(names, pretty), key=lambda t: -rows[t[0]]["regs"]):
        r, c = rows[n], counts.get(n, {k: 0 for k in MNEMONICS})
        flag = "!" if r["spill_st"] or r["spill_ld"] else " "
        print(f"{r['regs']:>5} {r['stack']:>6} {r['spill_st'] + r['spill_ld']:>6}{flag} {r['smem']:>6}  " +
              " ".join(f"{c[k]:>8}" for k in MNEMONICS) + f"  {p[:110]}")


if __name__ == "__main__":
    main()
