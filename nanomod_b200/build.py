"""Build nanomod_b200/_C/libnanomod_b200.so in-tree with nvcc for sm_100a.

Used by ``__graft_entry__.build()`` and by developers (``python -m nanomod_b200.build``).
The two translation units compile in parallel; an object is rebuilt only when one of its
sources is newer.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libnanomod_b200.so")
UNITS = ["nm_api.cu", "nm_lane_kernel.cu", "nm_deep_kernel.cu", "nm_huge.cu", "nm_rank.cu", "nm_downsample.cu", "nm_format.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# Default configuration of the lane tier: float32 sort keys, compare-exchanges mixed 1:2 between
# {FMNMX, FMNMX} (ALU pipe) and {FMNMX, a + b - min on the bit patterns as two IMADs} (FMA pipe).
# No key conversion at all; measured 0.8 % faster than the int32-key form (-DNM_INT_KEYS) of
# round 1 (profiles/round2_lane_variants.md).
# NM_WALK16_IMAD: the pointer bumps of the grid-key kernel's KS walk as IMADs (its ALU pipe is the busier one:
# 1.565 -> 1.522 ms, profiles/round2_grid_keys.md).
DEFAULT_DEFS = ("NM_FLOAT_IMAD", "NM_WALK16_IMAD")


def _deps():
    inc = os.path.join(HERE, "..", "include", "nanomod_b200.h")
    return [inc] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                    if f.endswith((".cuh", ".inc", ".h"))]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(unit, force, defs=(), suffix=""):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OUT_DIR, unit.replace(".cu", suffix + ".o"))
    if not force and not _stale(obj, [src] + _deps()):
        return obj, ""
    cmd = ["nvcc"] + NVCC_FLAGS + ["-D" + d for d in (defs or DEFAULT_DEFS)] + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False, defs=(), suffix: str = "") -> str:
    """Build the library.  ``defs``/``suffix`` make an experimental variant (e.g.
    defs=("NM_INT_KEYS",), suffix="_int" -> libnanomod_b200_int.so) that can be selected at run
    time with the NANOMOD_B200_LIB environment variable; the default build has neither."""
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = LIB.replace(".so", suffix + ".so")
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(lambda u: _compile(u, force, defs, suffix), UNITS))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    if log:
        with open(os.path.join(OUT_DIR, "ptxas%s.log" % suffix), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    if force or _stale(lib, objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    _defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    _suf = "".join(a[len("--suffix="):] for a in sys.argv[1:] if a.startswith("--suffix="))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defs=_defs, suffix=_suf))
