"""Builds libdqmc_b200.so (sm_100a only) in-tree with nvcc.  No CPU fallback exists."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libdqmc_b200.so"
SOURCES = ["gemm.cu", "udt.cu", "udt_reg.cu", "udt_steps.cu", "rdivp.cu", "update.cu", "update3.cu", "slicestep.cu", "misc.cu", "capi.cu", "ut.cu", "measure.cu", "global.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((HERE.parent / "include").glob("*.h"))
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    objs = []
    procs = []
    for src in SOURCES:
        obj = CSRC / (src[:-3] + ".o")
        cmd = ["nvcc", *flags, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libdqmc_b200.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(OUT), *objs,
                           "-lcudart", "-ldl"])
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
