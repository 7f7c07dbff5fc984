#!/usr/bin/env python
"""profiles/sass_tcgen05.txt: counts of the tensor-core / bulk-copy / barrier instructions in the SASS of the built library.

    python profiles/make_sass_tcgen05.py        (needs cuobjdump and leaf-grasping-vision-ml_b200/liblgb200.so)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = """SASS evidence for the tensor-core / bulk-copy path (cuobjdump -sass leaf-grasping-vision-ml_b200/liblgb200.so, sm_100a;
profiles/make_sass_tcgen05.py).  Counts of the instructions the B200 profiling recipe names, per kernel that contains a
tensor-core or bulk-copy instruction:
  UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,
  R2UR = vector-to-uniform register moves (143 per convolution kernel before the warp index was made provably uniform)
"""


def main():
    so = os.path.join(ROOT, "leaf-grasping-vision-ml_b200", "liblgb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"\b(UTCHMMA|LDTM|UTCBAR|UBLKCP|SYNCS|R2UR)\b")
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
        elif cur:
            mm = pat.search(line)
            if mm:
                cnt[cur][mm.group(1)] += 1
    with open(os.path.join(ROOT, "profiles", "sass_tcgen05.txt"), "w") as f:
        f.write(HDR + "\n")
        for k, c in cnt.items():
            if c["UTCHMMA"] or c["UBLKCP"] or c["LDTM"]:
                f.write(k[:100] + "\n    " + ", ".join(f"{n}: {c[n]}" for n in sorted(c)) + "\n")


if __name__ == "__main__":
    main()
