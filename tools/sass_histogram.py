"""Instruction histogram per kernel of libwaveglow_b200.so (cuobjdump -sass), for profiles/*_sass_summary.txt.

    python tools/sass_histogram.py > profiles/r02a_sass_summary.txt

One block per kernel: total SASS instructions and the counts of the mnemonics that prove what the kernel is built
from -- tcgen05 MMA (UTCHMMA / UTCQMMA...), TMEM loads (LDTM), TMEM alloc (UTCATOMSWS...), TMA (UTMALDG / UTMASTG /
UTMAPF), mbarrier traffic (SYNCS), tcgen05.commit (UTCBAR), MUFU (tanh / ex2), legacy tensor ops (HMMA: must be 0),
programmatic dependent launch (PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "text2speech_b200", "libwaveglow_b200.so")
INTEREST = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "LDTM", "STTM", "UTCATOMSWS", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG",
            "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "SYNCS", "MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "HMMA", "IMMA", "DMMA", "FFMA",
            "LDG", "STG", "LDS", "STS", "RED", "ATOM", "BAR.SYNC", "UCGABAR", "SHFL", "PREEXIT", "ACQBULK")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for key in INTEREST:
                if op == key or op.startswith(key + ".") or (key.startswith("MUFU") and op.startswith(key)):
                    kernels[cur][key] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in line:
                kernels[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and ".2CTA" in line:
                kernels[cur]["UTMALDG.2CTA"] += 1
            if op.startswith("UTMALDG"):
                d = re.search(r"UTMALDG\.(\dD)", line)
                if d:
                    kernels[cur]["UTMALDG." + d.group(1)] += 1
    names = list(kernels)
    try:
        dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    except (OSError, subprocess.CalledProcessError):
        pass
    print(f"# SASS instruction histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), one block per kernel")
    print("# tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, TMA = UTMALDG / UTMASTG / UTMAPF, "
          "mbarrier = SYNCS; HMMA / IMMA (mma.sync) must not appear in the GEMM kernels")
    tc = 0
    for k, c in kernels.items():
        name = demangle.get(k, k)
        name = re.sub(r">\([^<>]*\)$", ">", name) if ">(" in name else re.sub(r"\(.*\)$", "", name)
        keys = [x for x in c if x != "_total"]
        body = "  ".join(f"{x}={c[x]}" for x in sorted(keys))
        print(f"\n{name}\n    instructions={c['_total']}  {body}")
        tc += 1 if any(x.startswith("UTC") and x.endswith("MMA") or x == "UTCHMMA" for x in keys) else 0
    print(f"\n# {len(kernels)} kernels, {tc} of them issue tcgen05.mma")


if __name__ == "__main__":
    sys.exit(main())
