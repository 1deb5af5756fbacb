"""SASS opcode histogram per kernel of the shipped library (evidence that the hot path is tcgen05 / TMEM / TMA code:
UTCHMMA = tcgen05.mma, STTM / LDTM = tcgen05.st / ld, UTMALDG / UBLKCP = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
no HMMA = no legacy mma.sync).  usage: python tools/sass_histogram.py [lib.so] > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "quick_b200", "libquick_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "STTM", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "STAS", "UCGABAR", "ELECT", "HMMA", "LDGSTS", "MULTIMEM", "RED", "MEMBAR", "HFMA2", "HMUL2", "HADD2", "LOP3"]
cur, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        hist[cur][op] += 1
        full = line.split("*/", 1)[1] if "*/" in line else ""
        if "MULTIMEM" in full:
            hist[cur]["MULTIMEM"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} — opcode counts per kernel (base mnemonic, modifiers stripped)")
print("# " + " ".join(KEY) + " | total")
for (name, c), pretty in zip(hist.items(), demangle):
    short = re.sub(r"\(CUtensorMap_st, qb200::GemmArgs\)", "", pretty)
    short = re.sub(r"void |qb200::|\(anonymous namespace\)::", "", short)[:70]
    print(f"{short:70s} " + " ".join(f"{sum(v for kk, v in c.items() if kk.startswith(k)):4d}" for k in KEY) + f" | {sum(c.values())}")
tot = collections.Counter()
for c in hist.values():
    tot.update(c)
print(f"{'ALL KERNELS':70s} " + " ".join(f"{sum(v for kk, v in tot.items() if kk.startswith(k)):4d}" for k in KEY) + f" | {sum(tot.values())}")
