"""Times the UDT kernel alone (296 matrices of n = 256 by default) with CUDA events via the profile hooks."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _b200_loader
pkg = _b200_loader.load()
from oracle import model as OM
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
U = -4.0
T = OM.hopping_matrix("square", (L, L)); N, M = L * L, 20
e2, e2i, eh, ehi = OM.hopping_exponentials(T, 0.1)
ctx = pkg.Context(n_sites=N, n_slices=M, field_kind=1, n_chains=B, ranges=OM.generate_chunks(M, 10),
                  alpha=OM.hirsch_alpha(U, 0.1, 1), hopping_exp_squared=e2, hopping_exp_inv_squared=e2i,
                  hopping_exp=eh, hopping_exp_inv=ehi, seed=1)
g = np.random.default_rng(1)
ctx.set_conf(np.asfortranarray(g.choice(np.array([-1, 1], dtype=np.int8), size=(N, M, B))))
ctx.build_stack()
ctx.profile(True)
ctx.forward_build_stack()
rep = ctx.profile_report()
print({k: (round(v["ms"] / max(v["count"], 1), 3), v["count"]) for k, v in rep.items()})
