"""Times the update kernel class of a bench configuration over two sweeps: python profiles/time_update_cfg.py cfg5 [chains]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
import _b200_loader
pkg = _b200_loader.load()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg][4]
mc, _, _ = bench.make_mc(pkg, cfg, B, 0)
mc.ctx.build_stack()
mc.ctx.sweep(1)
mc.ctx.profile(True)
acc = mc.ctx.sweep(1)
rep = mc.ctx.profile_report()
print(cfg, B, "acceptance", float(acc.mean()) / (mc.ctx.n_sites * mc.ctx.n_slices * 2) if hasattr(mc.ctx, "n_sites") else acc.mean(),
      {k: (round(v["ms"] / max(v["count"], 1), 4), v["count"], round(v["ms"], 1)) for k, v in rep.items()})
