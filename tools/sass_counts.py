"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (tcgen05 = UTC*MMA, TMA = UTMALDG /
UTMASTG / UBLKCP, TMEM = LDTM / STTM) in the shipped library.  Writes profiles/r02_sass_counts.txt.

    python tools/sass_counts.py [path/to/libvatss_b200.so]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "speech_separation_b200", "libvatss_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU", "HMMA", "FFMA"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(CUtensorMap_st.*", "(...)", cur)
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in MNEMONICS:
            if op.startswith(k):
                counts[cur][k] += 1
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = [f"# SASS mnemonic counts per kernel of {os.path.relpath(so, ROOT)} (cuobjdump -sass, built from commit {head}+)",
       "# UTCHMMA = tcgen05.mma kind::f16, UTMALDG / UTMASTG / UBLKCP = TMA, LDTM / STTM = tcgen05.ld / tcgen05.st, HMMA = legacy mma.sync",
       f"{'kernel':90s} " + " ".join(f"{k:>8s}" for k in ["instr"] + MNEMONICS)]
for k in order:
    c = counts[k]
    out.append(f"{k[:90]:90s} " + " ".join(f"{c[m]:8d}" for m in ["instructions"] + MNEMONICS))
open(os.path.join(ROOT, "profiles", "r02_sass_counts.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:3] + [l for l in out[3:] if " k_tc_" in l or "k_tail" in l or "k_encoder" in l][:40]))
