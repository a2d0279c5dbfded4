"""One eager sweep of a configuration for ncu: python profiles/sweep_one.py cfg2 [chains]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
import _b200_loader
pkg = _b200_loader.load()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg][4]
mc, _, _ = bench.make_mc(pkg, cfg, B, 0)
mc.ctx.build_stack()
print("accepted", mc.ctx.sweep(1).mean(), "launches", mc.ctx.kernel_launches())
