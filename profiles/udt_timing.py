"""Times dqmc_op-level UDT calls on device-resident data through a Context's own sweep profile: prints per-class ms of one
sweep (K sweeps) for a configuration.  Usage: python profiles/udt_timing.py cfg4 [chains] [sweeps]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
import _b200_loader
pkg = _b200_loader.load()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg][4]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mc, _, _ = bench.make_mc(pkg, cfg, B, 0)
ctx = mc.ctx
ctx.build_stack()
ctx.sweep(1)
ctx.profile(True)
for _ in range(K):
    ctx.sweep(1)
prof = ctx.profile_report()
ctx.profile(False)
out = {k: {"ms_per_sweep": v["ms"] / K, "calls_per_sweep": v["count"] / K, "avg_ms": v["ms"] / max(v["count"], 1)} for k, v in prof.items()}
out["total_ms_per_sweep"] = sum(v["ms"] for v in prof.values()) / K
out["config"] = cfg; out["chains"] = B
print(json.dumps(out))
