"""montecarlo.jl_b200 -- B200-native DQMC sweep path behind MonteCarlo.jl's API.

The product is the CUDA library (csrc/ -> libdqmc_b200.so, C ABI in include/dqmc_b200.h);
this package is the thin host-side mirror of the reference's user API for that path.
Importing it never touches the CPU oracle; creating a context without a B200 raises.
"""
from . import _lib, build
from .context import (Context, DQMCError, FIELD_DENSITY_GHQ, FIELD_DENSITY_HIRSCH, FIELD_MAGNETIC_GHQ,
                      FIELD_MAGNETIC_HIRSCH, calculate_greens_AVX,
                      rdivp, udt_AVX_pivot, vmul)
from .dqmc import (ConfigRecorder, DQMC, DQMCParameters, DeviceMeasurement, Discarder, Field, GlobalFlip, GlobalShuffle, GreensMeasurement, HirschField,
                   LocalSweep, SimpleScheduler, charge_density_correlation, charge_density_susceptibility,
                   generate_chunks, greens_measurement, interaction_energy, kinetic_energy, occupation, replay, run, run_b,
                   spin_density_correlation, spin_density_susceptibility, sym_exp, total_energy)
from .lattices import Bond, Chain, Honeycomb, Lattice, SquareLattice, TriangularLattice, UnitCell
from .models import (HubbardModel, HubbardModelAttractive, HubbardModelRepulsive, choose_field, hopping_matrix,
                     lattice)

__all__ = [n for n in dir() if not n.startswith("_")]
