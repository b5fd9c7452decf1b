"""Build ``libhelp_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no torch)."""
from __future__ import annotations

import os
import os.path as osp
import shutil
import subprocess
import sys

HERE = osp.dirname(osp.abspath(__file__))
SRC = osp.join(HERE, "csrc")
OUT = osp.join(HERE, "libhelp_b200.so")
SOURCES = ["help_capi.cu"]


def deps() -> list:
    """Every source the library is built from: all of csrc/ plus the public header."""
    import glob
    return sorted(glob.glob(osp.join(SRC, "*.cu")) + glob.glob(osp.join(SRC, "*.cuh")) + glob.glob(osp.join(SRC, "*.h"))) + \
        [osp.join(HERE, "..", "include", "help_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # every REAL*4 / REAL*8 operation rounds once, as in the reference's SSE2 code;
    # contraction into FMA would move results away from the oracle (DESIGN.md "Numerics")
    "-fmad=false",
    # the simulation kernel is compiled in 13 instantiations of ~300 KB of SASS each: one ptxas job per kernel, in parallel
    "-split-compile", "0",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not osp.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build() -> bool:
    if not osp.exists(OUT):
        return True
    t = osp.getmtime(OUT)
    return any(osp.getmtime(d) > t for d in deps())


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """Compile the engine.  `out`/`defines` are for developer A/B builds (scripts/)."""
    if force or out or needs_build():
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
            ["-D" + d for d in defines] + ["-o", out or OUT] + [osp.join(SRC, s) for s in SOURCES]
        env = dict(os.environ)
        r = subprocess.run(cmd, cwd=SRC, env=env, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out or OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
