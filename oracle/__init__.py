"""ORACLE package (test infrastructure, NOT product code).

CPU restatement of the reference's spatial-feature-memory path:
  * ``oracle/geometry.c``      plain-C back-projection / quantisation / pooling / sequential cell sums
  * ``oracle/paste.c``         plain-C mask pasting (detectron2 paste_masks_in_image as called at custom_rcnn.py:880)
  * ``oracle/reference_ops.py`` torch-CPU restatement of the cited reference lines

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product package never does (tests/test_no_oracle_in_product.py
enforces it).

Parity pin: fixtures under ``tests/golden/`` were produced by running the reference's own source
(``tests/golden/make_golden.py``, needs /root/reference) and are checked against this oracle by
``tests/test_oracle_golden.py``.  The SMNet height-max tie rule (torch_scatter 1.4.0, absent) and the two
bytecode-only write variants are "parity unpinned" - see DESIGN.md.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("geometry.c", "paste.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def backproject_quantize(depth: np.ndarray, T: np.ndarray, intr, shift0, shift1, cell: float, map_w: int,
                         map_h: int, order: int = 0, z_clip: float = 0.5, want=("idx", "q2", "outlier", "height", "world")):
    """depth (H,W) f32; T (4,4) or (3,4) f32.  Returns dict of the requested outputs."""
    depth = np.ascontiguousarray(depth, np.float32)
    H, W = depth.shape
    T12 = np.ascontiguousarray(np.asarray(T, np.float32).reshape(-1)[:12])
    intr = np.ascontiguousarray(intr, np.float32)
    s0 = np.ascontiguousarray(shift0, np.float32)
    s1 = np.ascontiguousarray(shift1, np.float32)
    out = {
        "idx": np.empty((H, W), np.int32) if "idx" in want else None,
        "q2": np.empty((H, W, 2), np.int32) if "q2" in want else None,
        "outlier": np.empty((H, W), np.uint8) if "outlier" in want else None,
        "height": np.empty((H, W), np.float32) if "height" in want else None,
        "world": np.empty((H, W, 3), np.float32) if "world" in want else None,
    }
    lib().oracle_backproject_quantize(_p(depth), H, W, _p(T12), _p(intr), _p(s0), _p(s1), ctypes.c_float(cell),
                                      map_w, map_h, order, ctypes.c_float(z_clip), _p(out["idx"]), _p(out["q2"]),
                                      _p(out["outlier"]), _p(out["height"]), _p(out["world"]))
    return {k: v for k, v in out.items() if v is not None}


def read_pool_f16(table16: np.ndarray, idx: np.ndarray):
    """table16 (cells,C) float16, idx (H,W) int32 -> L0,L1,L2 float16 (C,h,w)."""
    table16 = np.ascontiguousarray(table16, np.float16)
    idx = np.ascontiguousarray(idx, np.int32)
    C = table16.shape[1]
    H, W = idx.shape
    L0 = np.empty((C, H // 8, W // 8), np.float16)
    L1 = np.empty((C, H // 16, W // 16), np.float16)
    L2 = np.empty((C, H // 32, W // 32), np.float16)
    lib().oracle_read_pool_f16(_p(table16.view(np.uint16)), C, _p(idx), H, W, _p(L0.view(np.uint16)),
                               _p(L1.view(np.uint16)), _p(L2.view(np.uint16)))
    return L0, L1, L2


def cell_sums_seq(feat_chw: np.ndarray, idx: np.ndarray, samp: Optional[np.ndarray], n_cells: int):
    """Sequential raster-order fp32 per-cell sums.  feat (C,H,W) f32 -> (sums (cells,C) f32, n (cells,) i32)."""
    feat = np.ascontiguousarray(feat_chw, np.float32)
    C = feat.shape[0]
    idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
    HW = idx.shape[0]
    s = np.zeros((n_cells, C), np.float32)
    n = np.zeros((n_cells,), np.int32)
    sm = None if samp is None else np.ascontiguousarray(samp, np.uint8).reshape(-1)
    lib().oracle_cell_sums_seq(_p(feat), C, HW, _p(idx), _p(sm), _p(s), _p(n))
    return s, n


def paste_masks(probs: np.ndarray, boxes: np.ndarray, H: int, W: int, thr: float = 0.5, want_values: bool = False, skip_empty: bool = True):
    """probs (K,S,S) f32, boxes (K,4) f32 XYXY -> masks (K,H,W) bool [, sampled values (K,H,W) f32] (oracle/paste.c).
    skip_empty=True: detectron2's CPU path (integer neighbourhood of the box); False: its CUDA path (whole image), which
    is what the library follows for thresholds below 0.5."""
    probs = np.ascontiguousarray(probs, np.float32)
    boxes = np.ascontiguousarray(boxes, np.float32)
    K, S = probs.shape[0], probs.shape[1]
    masks = np.zeros((K, H, W), np.uint8)
    values = np.zeros((K, H, W), np.float32) if want_values else None
    if K:
        lib().oracle_paste_masks2(_p(probs), _p(boxes), K, S, H, W, ctypes.c_float(thr), int(bool(skip_empty)), _p(masks), _p(values))
    return (masks.astype(bool), values) if want_values else masks.astype(bool)
