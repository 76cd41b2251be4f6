"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.ncu-summary.txt:

    python tools/ncu_summary.py gpurun_out/prof_gemm_pair.ncu-rep
"""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"^(Kernel Name|Block Size|Grid Size|Cluster|launch__cluster|dram__bytes_(read|write)\.sum|gpu__dram_throughput|gpu__time_duration\.sum|"
                 r"launch__(registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit|grid_size|block_size)|sm__cycles_elapsed\.max|"
                 r"sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct|"
                 r"lts__t_bytes\.sum($|\.per_second)|lts__throughput\.avg\.pct|lts__t_sector_hit_rate\.pct|l1tex__m_xbar2l1tex_read_bytes\.sum($|\.per_second)|"
                 r"smsp__average_warps_issue_stalled_.*_per_issue_active|TPC\.TriageCompute\.sm__pipe_tensor_cycles_active_realtime)")


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        for h, u, v in zip(hdr, units, vals):
            if PAT.match(h) and v != "":
                out.append("%-100s %-12s %s" % (h, u, v))
        out.append("")
    name = os.path.splitext(os.path.basename(rep))[0]
    path = os.path.join(ROOT, "profiles", name + ".ncu-summary.txt")
    open(path, "w").write("\n".join(out))
    print("\n".join(out))


if __name__ == "__main__":
    main()
