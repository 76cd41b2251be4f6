"""Per-kernel resource usage from ptxas (no GPU needed): registers, spills, static shared memory, barriers.
    python tools/ptxas_report.py  ->  profiles/ptxas_resources.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "distributed_tensorflow_b200", "csrc")
FILES = ["gemm_tcgen05.cu", "ps_engine.cu", "elementwise.cu", "step_exec.cu", "fabric_vmm.cu", "nn_kernels.cu"]


def main():
    rows = []
    for f in FILES:
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v",
                            "-c", os.path.join(CSRC, f), "-o", "/dev/null"], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stderr)
            raise SystemExit("nvcc failed on %s" % f)
        name = None
        for line in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and name:
                stack, st, ld = m.groups()
                rows.append([f, name, stack, st, ld, "", "", ""])
                continue
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?.*?(?:, (\d+) bytes smem)?", line)
            if m and name and rows and rows[-1][1] == name:
                regs = m.group(1)
                bars = re.search(r"used (\d+) barriers", line)
                smem = re.search(r"(\d+) bytes smem", line)
                rows[-1][5:] = [regs, bars.group(1) if bars else "0", smem.group(1) if smem else "0"]
    out = ["ptxas -v resource usage per kernel (sm_100a, -O3 -lineinfo; tools/ptxas_report.py)", "",
           "%-18s %-52s %5s %7s %12s %5s %9s" % ("file", "kernel", "regs", "stack B", "spill st/ld", "bars", "smem B")]
    for f, name, stack, st, ld, regs, bars, smem in rows:
        out.append("%-18s %-52s %5s %7s %12s %5s %9s" % (f, name[:52], regs, stack, "%s/%s" % (st, ld), bars, smem))
    spills = sum(int(r[3]) + int(r[4]) for r in rows)
    out += ["", "total spill bytes across all kernels: %d" % spills]
    path = os.path.join(ROOT, "profiles", "ptxas_resources.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
