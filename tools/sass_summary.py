"""SASS evidence per kernel of libl2i_b200.so: counts of the Blackwell-native mnemonics (tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM,
TMA = UTMALDG / UTMASTG / UBLKCP) and of the legacy tensor path (HMMA), plus registers from the ptxas logs.
    python tools/sass_summary.py > profiles/<tag>_sass_summary.txt        (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "latent2im_b200", "lib", "libl2i_b200.so")
PAT = {"UTC*MMA": r"\bUTC[A-Z]*MMA\b", "LDTM": r"\bLDTM\b", "UTMALDG": r"\bUTMALDG\b", "UTMASTG": r"\bUTMASTG\b", "UBLKCP": r"\bUBLKCP\b",
       "UTCBAR": r"\bUTCBAR\b", "LDGSTS": r"\bLDGSTS\b", "HMMA": r"\bHMMA\b", "ELECT": r"\bELECT\b", "R2UR": r"\bR2UR\b"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            counts[cur]["instructions"] += 1
            for k, p in PAT.items():
                if re.search(p, line):
                    counts[cur][k] += 1
    print(f"{os.path.relpath(LIB, ROOT)}: {len(counts)} kernels (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in PAT))
    tot = collections.Counter()
    for name, c in counts.items():
        full = demangle(name).replace("(anonymous namespace)::", "").replace("void ", "").replace("l2i::", "").replace("(int)", "").replace("(bool)", "")
        short = re.sub(r"\(CUtensorMap.*|\((float|void|__nv|unsigned|long|int|const).*", "", full)[:78]
        if c["UTC*MMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"] or c["UBLKCP"] or c["HMMA"] or "--all" in sys.argv:
            print(f"{short:78s} {c['instructions']:6d} " + " ".join(f"{c[k]:7d}" for k in PAT))
        tot.update(c)
    print(f"{'TOTAL (all kernels)':78s} {tot['instructions']:6d} " + " ".join(f"{tot[k]:7d}" for k in PAT))


if __name__ == "__main__":
    main()
