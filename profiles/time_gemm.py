"""Times the n x n x n batched GEMM classes alone via forward_build_stack (27 GEMM launches, 296 matrices)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _b200_loader
pkg = _b200_loader.load()
from oracle import model as OM
L, B = 16, 148
T = OM.hopping_matrix("square", (L, L)); N, M = L * L, 20
e2, e2i, eh, ehi = OM.hopping_exponentials(T, 0.1)
ctx = pkg.Context(n_sites=N, n_slices=M, field_kind=1, n_chains=B, ranges=OM.generate_chunks(M, 10),
                  alpha=OM.hirsch_alpha(-4.0, 0.1, 1), hopping_exp_squared=e2, hopping_exp_inv_squared=e2i,
                  hopping_exp=eh, hopping_exp_inv=ehi, seed=1)
g = np.random.default_rng(1)
ctx.set_conf(np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M, B))))
ctx.build_stack()
ctx.profile(True)
for _ in range(3):
    ctx.forward_build_stack()
rep = ctx.profile_report()
ms = rep["gemm"]["ms"] / rep["gemm"]["count"]
print("variant", os.environ.get("DQMC_EXP_GEMM", "0"), "gemm ms/launch", round(ms, 4), "TFLOP/s", round(2 * N**3 * B * 2 / ms / 1e9, 2))
