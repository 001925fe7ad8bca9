"""ctypes wrapper of ``oracle/nms_oracle.c`` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The C file restates the reference's greedy NMS (ssd3d.py:407-426, IoU of utils.py:105-149) for lists of millions
of boxes; it is compiled with gcc (``-ffp-contract=off``: every fp32 operation separately rounded, like the
reference's torch ops) into ``oracle/_build/libnms_oracle.so`` by ``build()`` (called from
``__graft_entry__.build()`` and lazily here).  Only ``tests/`` may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nms_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libnms_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(OUT_DIR, exist_ok=True)
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", LIB + ".tmp",
               SRC, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed for the NMS oracle:\n%s\n%s" % (r.stdout, r.stderr))
        os.replace(LIB + ".tmp", LIB)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.nms_oracle_greedy.restype = ctypes.c_longlong
        lib.nms_oracle_greedy.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p]
        _lib = lib
    return _lib


def greedy_nms(boxes_sorted, max_overlap: float) -> np.ndarray:
    """Keep mask (bool, n) of the reference's greedy NMS over score-sorted ``boxes_sorted`` ((n, 6) fp32 array or
    CPU tensor, [min, max] corners).  ``max_overlap`` is rounded to fp32 like torch does in ``iou > max_overlap``."""
    b = np.ascontiguousarray(np.asarray(boxes_sorted, dtype=np.float32))
    if b.ndim != 2 or b.shape[1] != 6:
        raise ValueError("boxes must be (n, 6)")
    keep = np.zeros((b.shape[0],), dtype=np.uint8)
    rc = _load().nms_oracle_greedy(b.ctypes.data, b.shape[0], float(np.float32(max_overlap)), keep.ctypes.data)
    if rc == -2:
        raise ValueError("nms_oracle: needs max_overlap >= 0 and finite boxes with max >= min")
    if rc < 0:
        raise MemoryError("nms_oracle: allocation failed")
    assert rc == int(keep.sum())
    return keep.astype(bool)
