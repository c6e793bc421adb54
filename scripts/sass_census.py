"""Per-kernel census of the Blackwell-native instructions in libunetb200.so (run after `make`):
  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor load / store,
  UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, REDG/RED = vector reductions into global memory.
    python scripts/sass_census.py > profiles/sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vickers_hardness_unet_b200", "libunetb200.so")
PAT = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "RED", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    cur, counts, order = None, collections.defaultdict(collections.Counter), []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for p in PAT:
                if op == p or op.startswith(p + "."):
                    counts[cur][p] += 1
            counts[cur]["_total"] += 1
    names = subprocess.run(["cu++filt", "-p"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
    print(f"# SASS census of {os.path.relpath(SO, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions per kernel")
    print(f"# {'kernel':70s} " + " ".join(f"{p:>8s}" for p in PAT) + "    total")
    for mangled, name in zip(order, names or order):
        short = name.replace("void ", "").replace("(int)", "")
        c = counts[mangled]
        print(f"{short[:72]:72s} " + " ".join(f"{c[p]:8d}" for p in PAT) + f" {c['_total']:8d}")
    tc = [n for n in order if counts[n]["UTCHMMA"]]
    print(f"# {len(tc)} of {len(order)} kernels issue tcgen05.mma (UTCHMMA); "
          f"{sum(1 for n in order if counts[n]['UTMALDG'])} use TMA loads, {sum(1 for n in order if counts[n]['UTMASTG'])} TMA stores; "
          f"legacy mma.sync (HMMA): {sum(counts[n]['HMMA'] for n in order)} instructions")


if __name__ == "__main__":
    sys.exit(main())
