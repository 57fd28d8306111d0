"""Aggregate an .ncu-rep source page per CUDA source line: share of stall samples, dominant stall
reasons, shared-memory wavefronts (actual vs ideal).  Usage: ncu_lines.py rep [launch_index] [top]"""
import csv
import io
import subprocess
import sys


def main(rep, launch=0, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    agg = {}
    name = ""
    for r in rows:
        if r and r[0] == "Function Name":
            name = r[1]
        if r and r[0] == "Line No":
            hdr = r
            idx = {}
            for i, h in enumerate(hdr):
                idx.setdefault(h, i)
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        line = int(r[0])
        a = agg.setdefault(line, {"src": r[1], "n": 0, "wf": 0, "wfi": 0, "inst": 0, "st": {}})
        if r[idx["Address"]] == "-":      # the CUDA line row carries the per-line totals
            a["n"] = int(r[idx["# Samples"]] or 0)
            a["wf"] = int(r[idx["L1 Wavefronts Shared"]] or 0)
            a["wfi"] = int(r[idx["L1 Wavefronts Shared Ideal"]] or 0)
            a["inst"] = int(r[idx["Instructions Executed"]] or 0)
            for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait", "stall_math", "stall_lg",
                      "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_dispatch", "stall_membar"):
                v = int(r[idx[k]] or 0)
                if v:
                    a["st"][k[6:]] = v
    tot = sum(a["n"] for a in agg.values()) or 1
    print(name[:100])
    print("total samples %d; shared wavefronts %d (ideal %d)" % (tot, sum(a["wf"] for a in agg.values()), sum(a["wfi"] for a in agg.values())))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = " ".join("%s=%d" % kv for kv in sorted(a["st"].items(), key=lambda kv: -kv[1])[:4])
        print("%5.1f%% L%-4d inst=%-8d wf=%d/%d | %-70s | %s" % (100.0 * a["n"] / tot, line, a["inst"], a["wf"], a["wfi"], a["src"].strip()[:70], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
