"""Short cfg4-shaped workload for ncu: n = 256, 2 flavor blocks, 148 chains (296 matrices per launch),
beta = 2 (M = 20, C = 2) so that one sweep is 8x shorter than cfg4 but every launch has cfg4's shape.
Usage: python profiles/prof_workload.py [nsweeps] [chains]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _b200_loader  # noqa: E402

pkg = _b200_loader.load()
from oracle import model as OM  # noqa: E402  (lattice -> hopping matrix inputs only)

nsweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
L = int(sys.argv[3]) if len(sys.argv) > 3 else 16
U = float(sys.argv[4]) if len(sys.argv) > 4 else -4.0
T = OM.hopping_matrix("square", (L, L))
N, M = L * L, 20
fk = OM.choose_field(U)
e2, e2i, eh, ehi = OM.hopping_exponentials(T, 0.1)
g = np.random.default_rng(1)
ctx = pkg.Context(n_sites=N, n_slices=M, field_kind=fk, n_chains=B, ranges=OM.generate_chunks(M, 10),
                  alpha=OM.hirsch_alpha(U, 0.1, fk), hopping_exp_squared=e2, hopping_exp_inv_squared=e2i,
                  hopping_exp=eh, hopping_exp_inv=ehi, seed=1)
ctx.set_conf(np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M, B))))
ctx.build_stack()
acc = ctx.sweep(nsweeps)
print("accepted/chain", acc.mean(), "launches", ctx.kernel_launches())
