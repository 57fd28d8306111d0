"""Static evidence from the built library (no GPU needed): registers / spills / shared memory per kernel from
`cuobjdump -res-usage`, and the SASS mnemonics that prove tcgen05 / TMEM / bulk-copy (TMA) use per kernel.
usage: python profiles/sass_evidence.py > profiles/rNN/sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tgcn_b200", "libtgcn_b200.so")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "LDGSTS",
             "STAS", "REDAS", "UCGABAR_ARV", "ACQBULK")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = " ".join(line.split())
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    names = sorted(usage)
    dm = demangle(names)
    print("%s: %d kernels (sm_100a)\n" % (os.path.basename(LIB), len(names)))
    for n in names:
        short = re.sub(r"\(.*", "", dm.get(n, n)).replace("void ", "").replace("tgcn::", "")
        mn = "  ".join("%s x%d" % kv for kv in sorted(counts[n].items()))
        print("%-58s %s%s" % (short[:58], usage[n], ("   | " + mn) if mn else ""))


if __name__ == "__main__":
    sys.exit(main())
