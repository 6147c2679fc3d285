#!/usr/bin/env python
"""Developer tool: SASS opcode histogram per kernel of the shipped library (the evidence file profiles/r02_sass_digest*.txt).
    python tools/sass_digest.py > profiles/r02_sass_digest_v2.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pointdsc_b200", "libpointdsc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "HMMA", "FFMA", "DFMA", "MUFU", "LDG", "STG", "STG.256", "LDS", "STS",
        "SHFL", "ATOMS", "ATOMG", "REDG"]
kern, cur, it = collections.OrderedDict(), None, iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", next(it))
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["total"] += 1
        base = op.split(".")[0]
        kern[cur][base] += 1
        if base == "STG" and ".256" in op:
            kern[cur]["STG.256"] += 1
print("# SASS opcode histogram per kernel of pointdsc_b200/libpointdsc_b200.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld / st = LDTM / STTM, cp.async.bulk (both directions) = UBLKCP, mbarrier = SYNCS,")
print("# cp.async = LDGSTS, st.global.v8.b32 = STG.256 (also counted under STG); no UTMALDG: operand tiles are contiguous images, no tensor map; no HMMA")
print("kernel | total | " + " | ".join(cols))
for k, c in sorted(kern.items(), key=lambda kv: -kv[1]["total"]):
    print(f"{k} | {c['total']} | " + " | ".join(str(c[x]) for x in cols))
