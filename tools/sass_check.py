#!/usr/bin/env python
"""Count the Blackwell-specific SASS mnemonics per kernel of libibm_b200.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_check.py > profiles/rNN_sass_check.md
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "inferbiomechanics_b200", "libibm_b200.so")
PAT = {"UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA", "LDTM/STTM (tcgen05.ld/st)": r"\b(LDTM|STTM)", "UTMALDG (TMA load)": r"\bUTMALDG",
       "UTMASTG/UTMAREDG (TMA store / reduce)": r"\b(UTMASTG|UTMAREDG)", "UBLKCP (bulk copy)": r"\bUBLKCP", "HMMA (mma.sync)": r"\bHMMA",
       "LDGSTS (cp.async)": r"\bLDGSTS", "LDSM (ldmatrix)": r"\bLDSM", "SYNCS (mbarrier)": r"\bSYNCS",
       "FADD2/FMUL2/FFMA2 (packed fp32)": r"\bF(ADD|MUL|FMA)2\b"}


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, rows = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern).replace("void ", "")
            rows[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        for name, pat in PAT.items():
            if re.search(pat, line):
                rows[kern][name] += 1
    print("# SASS check of libibm_b200.so (cuobjdump -sass, sm_100a): Blackwell-specific instructions per kernel\n")
    print("| kernel | " + " | ".join(PAT) + " |")
    print("|---|" + "---:|" * len(PAT))
    for k, c in rows.items():
        if sum(c.values()) == 0:
            continue
        print(f"| `{k[:90]}` | " + " | ".join(str(c[n]) if c[n] else "" for n in PAT) + " |")
    n_tc = sum(1 for c in rows.values() if c["UTC*MMA (tcgen05.mma)"])
    print(f"\n{len(rows)} kernels in the library; {n_tc} issue tcgen05.mma (UTC*MMA); none contains HGMMA/QGMMA (wgmma is sm_90a-only).")


if __name__ == "__main__":
    sys.exit(main())
