"""SASS evidence: per kernel of libsvit_sm100.so, counts of the Blackwell-native mnemonics (B200_PROFILING.md:
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR, packed fp32 ->
FFMA2) plus one sample line each.  Usage: python tools/sass_evidence.py > profiles/rN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "svit_b200", "libsvit_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTMAPF|FFMA2|FMUL2|FADD2|HMMA|SYNCS)\b")
cur, counts, sample = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        k = m.group(1)
        counts[cur][k] += 1
        sample.setdefault((cur, k), re.sub(r"\s+/\*.*$", "", line.strip()))
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}  (sm_100a); kernels with at least one Blackwell-native mnemonic")
tot = collections.Counter()
for fn, c in counts.items():
    if not c:
        continue
    tot.update(c)
    print(f"\n{fn}")
    print("    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR"):
        if (fn, k) in sample:
            print(f"      e.g. {sample[(fn, k)]}")
print("\n# totals: " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
print(f"# kernels in the library: {len(counts)}")
