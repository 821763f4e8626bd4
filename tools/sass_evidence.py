#!/usr/bin/env python
"""Counts the SASS mnemonics that show a kernel is Blackwell-native (tcgen05 -> UTC*MMA, tensor memory -> LDTM/STTM,
TMA -> UTMALDG, cluster barriers, ...) per kernel of librbod.so.  Runs without a GPU (cuobjdump only).

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "retrieval_based_object_detection_b200", "librbod.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA[\w.]*|LDTM[\w.]*|STTM[\w.]*|UTMALDG[\w.]*|UTMASTG[\w.]*|UBLKCP[\w.]*|HMMA[\w.]*|"
                 r"UTCBAR[\w.]*|SYNCS[\w.]*|LDGSTS[\w.]*|DFMA|DADD|REDUX[\w.]*|UCGABAR[\w.]*|HGMMA[\w.]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cnt, cur = collections.defaultdict(collections.Counter), None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for t in PAT.findall(ln):
                cnt[cur][t] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: occurrences of Blackwell-specific (and fp64 / legacy-MMA) mnemonics per kernel")
    for mangled, name in sorted(zip(cnt, names), key=lambda p: p[1]):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
        print(f"{short}")
        for k, v in sorted(cnt[mangled].items()):
            print(f"    {k:40s} {v}")
    total = collections.Counter()
    for c in cnt.values():
        total.update(c)
    print("# legacy tensor path (HMMA / HGMMA) anywhere:", sum(v for k, v in total.items() if k.startswith(("HMMA", "HGMMA"))))


if __name__ == "__main__":
    sys.exit(main())
