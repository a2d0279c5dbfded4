"""One batched udt_AVX_pivot! call (graded columns like a B-chain) for ncu: python profiles/udt_only.py n batch"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import _b200_loader
pkg = _b200_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = int(sys.argv[2]) if len(sys.argv) > 2 else 296
g = np.random.default_rng(0)
X = g.random((n, n, b)) * np.exp(3 * g.normal(size=(1, n, b)))
for _ in range(2):
    U, D, T, piv = pkg.udt_AVX_pivot(X)
print("ok", float(np.abs(D).max()))
