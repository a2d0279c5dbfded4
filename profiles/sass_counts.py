"""Opcode counts per kernel of the in-tree library (cuobjdump -sass): python profiles/sass_counts.py > profiles/sass_counts.txt"""
import collections, re, subprocess, sys
from pathlib import Path
so = Path(__file__).resolve().parent.parent / "montecarlo.jl_b200" / "libdqmc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
WATCH = ["DMMA", "DFMA", "DMUL", "DADD", "UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "LDGSTS", "SYNCS",
         "UCGABAR_ARV", "UCGABAR_WAIT", "MEMBAR", "CCTL", "CREDUX", "SHFL", "BAR", "LDS", "STS", "LDG", "STG", "MUFU", "VOTE"]
cur = None; counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
print(f"# {so.name}: SASS opcode counts per kernel (sm_100a).  tcgen05 (UTC*MMA, LDTM/STTM) is absent by necessity: it has no f64 kind;")
print("# FP64 tensor math is DMMA.  TMA: UTMALDG (tensor-map loads, gemm.cu), UBLKCP (bulk copies, udt_steps.cu); SYNCS = mbarrier ops.")
tot = collections.Counter()
for fn, c in counts.items():
    short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", short)
    items = " ".join(f"{k}={v}" for k, v in c.items() if k != "total" and v)
    print(f"{short[:70]:70s} total={c['total']:6d} {items}")
    tot.update(c)
print("ALL KERNELS".ljust(70), " ".join(f"{k}={v}" for k, v in tot.items()))
