"""Builds + loads the TEST-ONLY host emulation of the engine's warp code.

Same sources as the product (grok_alpha_zero_b200/csrc) compiled by g++ with -DGAZ_EMUL
(cooperative width 1 instead of a 32-lane warp).  It lets `-m "not gpu"` tests exercise the
exact tree / game logic on a CPU box.  The package never loads it.
"""
import os
import subprocess

from grok_alpha_zero_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(ROOT, "grok_alpha_zero_b200", "csrc")
OUT = os.path.join(HERE, "_emul", "libgaz_emul.so")
_cached = None


def load():
    global _cached
    if _cached is not None:
        return _cached
    srcs = [os.path.join(SRC, f) for f in ("gaz_engine.cu", "gaz_core.cuh")]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs):
        subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-DGAZ_EMUL", "-O2", "-fPIC", "-shared",
                               "-ffp-contract=off", "-fno-fast-math", "-o", OUT, srcs[0], "-lm"])
    _cached = _lib.bind(OUT)
    return _cached
