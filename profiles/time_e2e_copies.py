"""Times the host <-> device legs of the end-to-end sweep separately (pinned buffers, cfg 4 shape):
python profiles/time_e2e_copies.py [config] [chains]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
import _b200_loader
pkg = _b200_loader.load()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg][4]
mc, _, _ = bench.make_mc(pkg, cfg, B, 0)
ctx = mc.ctx
N, M, nb = ctx.N, ctx.M, ctx.nb
ctx.build_stack()
u = torch.rand((B, 2 * M, N), dtype=torch.float64).pin_memory()
g = torch.empty((B, nb, N, N), dtype=torch.float64).pin_memory()
c = torch.empty((B, M, N), dtype=torch.int8).pin_memory()
ctx.sweep(1, uniforms=u.numpy()); ctx.greens(out=g.numpy()); ctx.get_conf(out=c.numpy())
def t(f, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("sweep, device RNG      ms", round(t(lambda: ctx.sweep(1)), 2))
print("sweep, host uniforms   ms", round(t(lambda: ctx.sweep(1, uniforms=u.numpy())), 2), " (H2D", u.numel() * 8 / 1e6, "MB)")
print("greens D2H             ms", round(t(lambda: ctx.greens(out=g.numpy())), 2), " (", g.numel() * 8 / 1e6, "MB)")
print("conf D2H               ms", round(t(lambda: ctx.get_conf(out=c.numpy())), 2), " (", c.numel() / 1e6, "MB)")
