"""Times the update kernel (cfg4 shape: n = 256, 2 blocks, 148 chains) over one sweep of a beta = 2 chain."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _b200_loader
pkg = _b200_loader.load()
from oracle import model as OM
L, B = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 148
T = OM.hopping_matrix("square", (L, L)); N, M = L * L, 20
e2, e2i, eh, ehi = OM.hopping_exponentials(T, 0.1)
ctx = pkg.Context(n_sites=N, n_slices=M, field_kind=1, n_chains=B, ranges=OM.generate_chunks(M, 10),
                  alpha=OM.hirsch_alpha(-4.0, 0.1, 1), hopping_exp_squared=e2, hopping_exp_inv_squared=e2i,
                  hopping_exp=eh, hopping_exp_inv=ehi, seed=1)
g = np.random.default_rng(1)
ctx.set_conf(np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M, B))))
ctx.build_stack()
ctx.sweep(1)
ctx.profile(True)
ctx.sweep(2)
rep = ctx.profile_report()
print("stagger", os.environ.get("DQMC_UPD_STAGGER_NS", "0"), os.environ.get("DQMC_UPD_STAGGER_CLASSES", "2"),
      {k: round(v["ms"] / max(v["count"], 1), 4) for k, v in rep.items()})
