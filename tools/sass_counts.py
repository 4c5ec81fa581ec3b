#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS mnemonics in the shipped library
(cuobjdump -sass cerberus_b200/libcerberus_b200.so): UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA
load / store), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier). Evidence that every
convolution kernel runs on tcgen05 + TMA; written to profiles/ by `python tools/sass_counts.py`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cerberus_b200", "libcerberus_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS",
       "HMMA", "IMMA", "FFMA", "HFMA2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "").replace("cerb::", "")
            cur = re.sub(r"^void ", "", cur)
            cur = re.sub(r"\((?!.*>).*$", "", cur) if ">" in cur else re.sub(r"\(.*$", "", cur)
            counts[cur] = collections.Counter()
            counts[cur]["_instr"] = 0
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_instr"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    lines = ["# cuobjdump -sass cerberus_b200/libcerberus_b200.so   (architectures in the fatbin: %s)" % ", ".join(arch),
             "# per kernel: SASS instruction count and occurrences of the tensor-core / TMA / TMEM mnemonics",
             "%-58s %7s " % ("kernel", "instr") + " ".join("%8s" % p for p in PAT)]
    for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["_instr"]):
        lines.append("%-58s %7d " % (k[:58], c["_instr"]) + " ".join("%8d" % c[p] for p in PAT))
    text = "\n".join(lines) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_counts.txt")
    open(dst, "w").write(text)
    print(text[:3000])


if __name__ == "__main__":
    main()
