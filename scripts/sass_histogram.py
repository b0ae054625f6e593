"""Instruction counts per kernel of libnsf_b200.so (cuobjdump -sass) -> profiles/r2_sass_histogram.txt
usage: python scripts/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nsfnet_b200/libnsf_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
cols = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "REDG", "RED.", "FFMA", "MUFU", "SHFL", "STS", "LDS", "LDGSTS", "R2UR", "LDL", "STL"]
print("instruction counts per kernel of nsfnet_b200/libnsf_b200.so (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,")
print("UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA), LDGSTS = cp.async, SYNCS = mbarrier ops, REDG = red.global, LDL / STL = local memory (spills)\n")
print(f"{'kernel':70s}" + "".join(f"{c:>9s}" for c in cols) + f"{'total':>9s}")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
rows = []
for nm, blk in zip(names, blocks):
    ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", blk)
    short = re.sub(r"\(.*", "", nm.replace("(anonymous namespace)::", "").replace("void ", ""))
    rows.append((short, [sum(1 for i in ins if i.startswith(c)) for c in cols], len(ins)))
for short, cnt, tot in sorted(rows):
    print(f"{short[:70]:70s}" + "".join(f"{c:9d}" for c in cnt) + f"{tot:9d}")
