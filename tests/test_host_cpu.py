"""CPU-side checks of the product: the C-ABI library loads and exports every declared symbol,
the host mirror of the reference API agrees with the oracle's restatement and the reference's
golden vectors, and the product refuses to run without a GPU (no CPU fallback).  No compute calls."""
import ctypes
import inspect
from pathlib import Path

import numpy as np
import pytest

from oracle import model as OM

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(b200):
    lib = b200._lib.load()
    names = b200._lib.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    raw = ctypes.CDLL(str(b200._lib.library_path()))
    for n in names:
        getattr(raw, n)


def test_library_is_sm100a_only(b200):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(b200._lib.library_path())], capture_output=True, text=True).stdout
    archs = {l.split("sm_")[1].split(".")[0] for l in out.splitlines() if "sm_" in l}
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_and_no_oracle_in_product(b200):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = b200.HubbardModel(b200.SquareLattice(4), U=1.0)
    with pytest.raises(b200.DQMCError, match="no CUDA device"):
        b200.DQMC(m, beta=1.0)
    # the product must not import / link the oracle
    for f in (ROOT / "montecarlo.jl_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh"):
            txt = f.read_text()
            assert "oracle" not in txt.replace("no CPU oracle", "").lower() or f.name == "__init__.py", f
    assert b200._lib.load().dqmc_max_sites() >= 288


def test_lattice_golden_bonds(b200):
    """test/lattices.jl:80-93, 169-182"""
    l = b200.SquareLattice(3)
    bs = l.bonds()
    assert [b.frm for b in bs] == [1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9]
    assert [b.to for b in bs] == [2, 4, 3, 5, 1, 6, 5, 7, 6, 8, 4, 9, 8, 1, 9, 2, 7, 3]
    bs = l.bonds(directed=True)
    assert [b.to for b in bs] == [2, 4, 3, 7, 3, 5, 1, 8, 1, 6, 2, 9, 5, 7, 6, 1, 6, 8, 4, 2, 4, 9, 5, 3,
                                  8, 1, 9, 4, 9, 2, 7, 5, 7, 3, 8, 6]
    h = b200.Honeycomb(2)
    assert len(h) == 8
    bs = h.bonds()
    assert [b.frm for b in bs] == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4]
    assert [b.to for b in bs] == [5, 6, 7, 6, 5, 8, 7, 8, 5, 8, 7, 6]
    bs = h.bonds(directed=True)
    assert [b.frm for b in bs] == [1, 1, 1, 5, 5, 5, 2, 2, 2, 6, 6, 6, 3, 3, 3, 7, 7, 7, 4, 4, 4, 8, 8, 8]
    assert [b.to for b in bs] == [5, 6, 7, 1, 2, 3, 6, 5, 8, 2, 1, 4, 7, 8, 5, 3, 4, 1, 8, 7, 6, 4, 3, 2]


def test_hopping_matrix_agrees_with_oracle(b200):
    for ctor, kind, Ls in ((b200.SquareLattice, "square", (4, 4)), (b200.Honeycomb, "honeycomb", (3, 3)),
                           (b200.Chain, "chain", (8,)), (b200.TriangularLattice, "triangular", (4, 4))):
        m = b200.HubbardModel(ctor(*Ls[:1]), U=2.0, mu=0.3, t=0.7)
        assert np.array_equal(b200.hopping_matrix(m), OM.hopping_matrix(kind, Ls, t=0.7, mu=0.3))


def test_parameters_and_chunks(b200):
    """test/flavortests_DQMC.jl:4-18, 244-262"""
    P = b200.DQMCParameters
    p = P(beta=5.0); assert (p.beta, p.delta_tau, p.slices) == (5.0, 0.1, 50)
    p = P(beta=5.0, delta_tau=0.01); assert p.slices == 500
    p = P(beta=50.0, slices=20); assert p.delta_tau == 2.5
    p = P(delta_tau=0.1, slices=50); assert p.beta == 5.0
    with pytest.raises(ValueError):
        P(beta=5.0, delta_tau=0.1, slices=49)
    with pytest.raises(ValueError):
        P(safe_mult=3)
    with pytest.raises(NotImplementedError):
        P(beta=1.0, checkerboard=True)
    for M, s in ((160, 10), (5, 3), (23, 10), (100, 7)):
        assert b200.generate_chunks(M, s) == OM.generate_chunks(M, s)
    assert b200.choose_field(b200.HubbardModel(L=4, U=-1.0)) == "MagneticHirschField"
    assert b200.HubbardModelRepulsive(L=4, U=2.0).U == -2.0
    assert np.allclose(b200.sym_exp(np.zeros((3, 3))), np.eye(3))


def test_header_documents_reference_interfaces():
    h = (ROOT / "include" / "dqmc_b200.h").read_text()
    for cite in ("stack.jl", "local_updates.jl", "fields.jl", "greens.jl", "UDT.jl", "real.jl"):
        assert cite in h


def test_bravais_srctrg2dir_agrees_with_oracle(b200):
    """lattices/lattice_cache.jl:224-240"""
    from oracle import measure as OMS
    for ctor, Ls in ((b200.SquareLattice, (4, 4)), (b200.Honeycomb, (3, 3)), (b200.Chain, (6,))):
        l = ctor(*Ls[:1])
        assert np.array_equal(np.array(l.bravais_srctrg2dir()), OMS.bravais_srctrg2dir(Ls))
    l = b200.Lattice(b200.SquareLattice(2).unitcell, (3, 4))
    assert np.array_equal(np.array(l.bravais_srctrg2dir()), OMS.bravais_srctrg2dir((3, 4)))


def test_simple_scheduler(b200):
    """updates/scheduler.jl:236-289: cycles through the updates, LocalSweep(N) expands, local updates are required."""
    with pytest.raises(ValueError):
        b200.SimpleScheduler(b200.GlobalFlip())
    s = b200.SimpleScheduler(b200.LocalSweep(), b200.GlobalFlip(), b200.LocalSweep(2))
    kinds = [type(s.next()).__name__ for _ in range(8)]
    assert kinds == ["LocalSweep", "GlobalFlip", "LocalSweep", "LocalSweep"] * 2


def test_conf_compress_is_julia_bitarray_layout(b200):
    """fields.jl:331-334: BitArray(conf .== 1); a BitArray stores bit i of the column-major array in chunks[i >> 6]
    at position i & 63.  Round trip through decompress."""
    m = b200.HubbardModel(b200.SquareLattice(3), U=1.0)
    p = b200.DQMCParameters(beta=0.9)
    f = b200.HirschField("DensityHirschField", p, m, 2)
    f.rand(np.random.default_rng(3))
    ch = f.compress(chain=1)
    flat = f.confs[:, :, 1].ravel(order="F")
    assert ch.dtype == np.uint64 and len(ch) == (flat.size + 63) // 64
    for i, v in enumerate(flat):
        assert ((int(ch[i >> 6]) >> (i & 63)) & 1) == int(v == 1)
    old = f.confs[:, :, 1].copy()
    f.confs[:, :, 1] = 1
    f.decompress(ch, chain=1)
    assert np.array_equal(f.confs[:, :, 1], old)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses bench.py's stdout: one JSON line, whatever the libraries on the way print (NCCL announces its version
    on fd 1).  The reference arm runs on the CPU, so the contract is checked here on a tiny anchor configuration."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--config", "anchor6a", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["e2e"]["h2d_bytes_per_step"] == 0
