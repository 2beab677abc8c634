#!/usr/bin/env python
"""Top stalled SASS instructions of one ncu --set full capture (needs --import-source on, -lineinfo).

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][1][:160] if rows and len(rows[0]) > 1 else "")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    out, agg = [], {}
    for r in rows[hi + 1:]:
        try:
            n = int(r[idx["# Samples"]])
        except Exception:
            continue
        st = {h: int(r[idx[h]] or 0) for h in stallcols}
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
        out.append((n, r[idx["Address"]][-5:], r[idx["Source"]][:100], sorted(st.items(), key=lambda kv: -kv[1])[:2]))
    tot = sum(o[0] for o in out)
    print("total samples", tot)
    print("by reason:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for o in sorted(out, key=lambda o: -o[0])[:top]:
        print(f"{o[0]:6d} {100.0 * o[0] / tot:5.1f}% {o[1]} {o[2]}  {o[3]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
