"""A few slice steps of a configuration for ncu: python profiles/sweep_few.py cfg4 [chains] [nprop]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
import _b200_loader
pkg = _b200_loader.load()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg][4]
nprop = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mc, _, _ = bench.make_mc(pkg, cfg, B, 0)
ctx = mc.ctx
ctx.build_stack()
for _ in range(nprop):
    ctx.sweep_spatial()
    ctx.propagate(1)
print("ok")
