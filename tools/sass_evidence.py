"""SASS evidence (runs without a GPU): per-kernel counts of the mnemonics that prove the Blackwell path.

    python tools/sass_evidence.py  -> profiles/sass_mnemonics.txt

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA load (.2CTA = pair variant whose
completion bytes land on the leader CTA's mbarrier), UTCBAR = tcgen05.commit (.2CTA.MULTICAST = arrive on both CTAs
of the pair), SYNCS = mbarrier, LDGMC = multimem.ld_reduce (in-switch reduction over NVLink/NVSwitch),
*.STRONG.SYS / MEMBAR.*.SYS / REDG...SYS = system-scope peer-memory traffic and signalling.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "distributed_tensorflow_b200", "_lib", "libdtf_kernels.so")
KEEP = re.compile(r"^(UTCHMMA|UTCBAR|UTMALDG|UTMAPF|UTMASTG|UBLKCP|LDTM|LDGMC|SYNCS|MEMBAR|REDG|ATOMG|NANOSLEEP|HMMA|UCGABAR|UTCATOMSWS|"
                  r"ACQBULK|PREEXIT|"
                  r"LDG\.E\.[0-9.]*STRONG\.SYS|STG\.E\.[0-9.]*STRONG\.SYS|LD\.E\.[0-9.]*STRONG\.SYS|ST\.E\.[0-9.]*STRONG\.SYS)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    full = collections.OrderedDict()            # kernel -> its complete SASS listing (profiles/sass/<kernel>.sass)
    cur_lines = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = per.setdefault(name, collections.Counter())
            cur_lines = full.setdefault(name, [])
            cur_lines.append(line)
            continue
        if cur_lines is not None:
            cur_lines.append(line)
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            if KEEP.match(op):
                if op.startswith("SYNCS"):
                    op = ".".join(op.split(".")[:2])
                cur[op] += 1
    out = [__doc__.strip().split("\n\n", 1)[1] if False else
           "SASS mnemonic counts per kernel in libdtf_kernels.so (cuobjdump -sass; sm_100a; tools/sass_evidence.py)",
           "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA load, UTCBAR = tcgen05.commit,",
           "SYNCS.* = mbarrier, LDGMC = multimem.ld_reduce (NVLS in-switch reduction), multimem.st = STG.E.*.STRONG.SYS on a",
           "multicast address, *.STRONG.SYS / MEMBAR.*.SYS / REDG..SYS = system-scope peer-memory signalling.",
           "HMMA (legacy mma.sync) total: %d" % sum(c.get("HMMA", 0) for c in per.values()), ""]
    for name, c in per.items():
        if not c:
            continue
        out.append(name)
        out.append("   " + ", ".join("%s x%d" % (k, v) for k, v in sorted(c.items())))
    sdir = os.path.join(ROOT, "profiles", "sass")
    os.makedirs(sdir, exist_ok=True)
    for old in os.listdir(sdir):
        os.unlink(os.path.join(sdir, old))
    for name, lines in full.items():
        fn = re.sub(r"[^A-Za-z0-9_]+", "_", name.replace("dtf::", "")).strip("_")[:120] + ".sass"
        # instruction text only (the /*hex encoding*/ columns double the size and carry no extra evidence)
        body = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", l) for l in lines if not re.match(r"^\s*/\* 0x[0-9a-f]{16} \*/\s*$", l)]
        with open(os.path.join(sdir, fn), "a") as f:
            f.write("\n".join(body) + "\n")
    print("wrote %d full listings to profiles/sass/" % len(full))
    path = os.path.join(ROOT, "profiles", "sass_mnemonics.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:6]))
    print("wrote", path, "(%d kernels)" % sum(1 for c in per.values() if c))


if __name__ == "__main__":
    sys.exit(main())
