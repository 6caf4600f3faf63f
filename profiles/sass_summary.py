#!/usr/bin/env python
"""SASS evidence that the shipped library is tcgen05 / TMEM / TMA code (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): per kernel of libsnn_heads_b200.so, the count of UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld),
UTMALDG (cp.async.bulk.tensor), UTCBAR (tcgen05.commit), SYNCS (mbarrier) and of the legacy HMMA / HGMMA mnemonics
(which must be zero).  Runs on the CPU box: python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "snn_automotive_object_detection_b200", "libsnn_heads_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\S*\.2CTA"), ("UTC*MMA (any)", r"\bUTC[A-Z]*MMA"),
            ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.2CTA", r"\bUTMALDG\S*\.2CTA"),
            ("UTCBAR", r"\bUTCBAR"), ("UTCBAR.MULTICAST", r"\bUTCBAR\S*MULTICAST"), ("SYNCS", r"\bSYNCS"),
            ("HMMA (legacy mma.sync)", r"\bHMMA"), ("HGMMA (Hopper wgmma)", r"\bHGMMA"),
            ("FSETP", r"\bFSETP"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("instructions", r"^\s+/\*[0-9a-f]{4}\*/")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                per[cur][name] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (arch: {sorted(set(arch))})")
    tot = collections.Counter()
    for fn, c in per.items():
        short = re.sub(r"\(.*", "", fn)
        print(f"\n{short}")
        print("   " + ", ".join(f"{k} {c[k]}" for k, _ in PATTERNS if c[k]))
        tot.update(c)
    print("\n# whole library")
    print("   " + ", ".join(f"{k} {tot[k]}" for k, _ in PATTERNS))
    assert tot["HMMA (legacy mma.sync)"] == 0 and tot["HGMMA (Hopper wgmma)"] == 0
    assert tot["UTC*MMA (any)"] > 0 and tot["LDTM"] > 0 and tot["UTMALDG"] > 0


if __name__ == "__main__":
    sys.exit(main())
