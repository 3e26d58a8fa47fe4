"""ctypes front end of the CPU oracle (oracle/pt_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  It reuses the product's *host-side* flatten and parameter structs (plain data
marshalling) but none of its compute: every number it returns is computed by pt_oracle.c in fp64,
sequentially, in the reference's operation order.  Parity pinned by tests/test_oracle_golden.py
against outputs of the unmodified reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from pathlib import Path
from typing import Optional

import numpy as np

from pytracer_b200 import _abi

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "pt_oracle.c"
_LIB = _HERE / "liboracle.so"
_lock = threading.Lock()
_lib = None


def build(force: bool = False) -> Path:
    """gcc -O2 -ffp-contract=off (no fused multiply-add: Python never fuses)."""
    header = _HERE.parent / "include" / "rt_api.h"  # the structs the oracle shares with the library
    newest = max(_SRC.stat().st_mtime, header.stat().st_mtime if header.exists() else 0.0)
    if force or not _LIB.exists() or _LIB.stat().st_mtime < newest:
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(_LIB), str(_SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            _lib = C.CDLL(str(build()))
            for name in ("orc_render", "orc_trace_rays", "orc_intersect", "orc_is_point_visible",
                         "orc_camera_rays", "orc_pcg_seed", "orc_pcg_draw", "orc_pigment_color",
                         "orc_scatter", "orc_onb"):
                getattr(_lib, name).restype = C.c_int
    return _lib


def _p(arr: Optional[np.ndarray]):
    return None if arr is None else C.c_void_p(arr.ctypes.data)


def render(flat, params: _abi.rt_render_params, row_begin: int = 0, row_end: Optional[int] = None,
           want_hit: bool = True, want_states: bool = False, out: Optional[np.ndarray] = None, row_step: int = 1):
    """Sequential fire_all_rays over rows [row_begin, row_end).  Returns a dict with rgb (H,W,3 f64),
    hit_index, counters (closest, shadow, samples), final aa/pt states and per-sample start states."""
    H, W = params.height, params.width
    row_end = H if row_end is None else row_end
    rgb = np.zeros((H, W, 3), dtype=np.float64) if out is None else out
    hit = np.full((H, W), -2, dtype=np.int32) if want_hit else None
    counters = np.zeros(3, dtype=np.uint64)
    aa = np.array([params.aa_state, params.aa_inc], dtype=np.uint64)
    pt = np.array([params.pt_state, params.pt_inc], dtype=np.uint64)
    spp = max(1, params.samples_per_side) ** 2
    n_rows = len(range(row_begin, row_end, row_step))
    states = np.zeros(n_rows * W * spp, dtype=np.uint64) if want_states else None
    rc = lib().orc_render(C.byref(flat.desc), C.byref(params), C.c_int(row_begin), C.c_int(row_end), C.c_int(row_step),
                          _p(rgb), _p(hit), _p(counters), _p(aa), _p(pt), _p(states))
    assert rc == 0
    return dict(rgb=rgb, hit_index=hit, rays_closest=int(counters[0]), rays_shadow=int(counters[1]),
                samples=int(counters[2]), aa_state=int(aa[0]), pt_state=int(pt[0]), sample_states=states)


def render_threaded(flat, params: _abi.rt_render_params, n_threads: int):
    """CPU-baseline helper: interleaved rows in parallel threads (ctypes drops the GIL).  Each thread
    gets its own jitter/scatter streams, so the image is a different but equally distributed draw; used
    for timing and for statistical references, never for bit-exact checks."""
    from pytracer_b200.pcg import PCG

    H, W = params.height, params.width
    rgb = np.zeros((H, W, 3), dtype=np.float64)
    results = [None] * n_threads

    def work(i):
        p = _abi.rt_render_params.from_buffer_copy(bytes(params))
        aa, pt = PCG(params.aa_state & 0xFFFFFFFF, 1000 + i), PCG(params.pt_state & 0xFFFFFFFF, 2000 + i)
        p.aa_state, p.aa_inc, p.pt_state, p.pt_inc = aa.state, aa.inc, pt.state, pt.inc
        results[i] = render(flat, p, i, H, want_hit=False, out=rgb, row_step=n_threads)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return dict(rgb=rgb, rays_closest=sum(r["rays_closest"] for r in results),
                rays_shadow=sum(r["rays_shadow"] for r in results), samples=sum(r["samples"] for r in results))


def trace_rays(flat, params, rays: np.ndarray, depth: Optional[np.ndarray] = None, pcg_state_inc=None):
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 8)
    n = rays.shape[0]
    depth_arr = None if depth is None else np.ascontiguousarray(depth, dtype=np.int32)
    st = np.array(pcg_state_inc if pcg_state_inc is not None else [params.pt_state, params.pt_inc], dtype=np.uint64)
    out = np.zeros((n, 3), dtype=np.float64)
    counters = np.zeros(3, dtype=np.uint64)
    rc = lib().orc_trace_rays(C.byref(flat.desc), C.byref(params), _p(rays), _p(depth_arr), C.c_int(n), _p(st), _p(out), _p(counters))
    assert rc == 0
    return out, (int(st[0]), int(st[1])), counters


def intersect(flat, rays: np.ndarray):
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 8)
    out = (_abi.rt_hit * rays.shape[0])()
    assert lib().orc_intersect(C.byref(flat.desc), _p(rays), C.c_int(rays.shape[0]), out) == 0
    return out


def is_point_visible(flat, pairs: np.ndarray) -> np.ndarray:
    pairs = np.ascontiguousarray(pairs, dtype=np.float64).reshape(-1, 6)
    out = np.zeros(pairs.shape[0], dtype=np.uint8)
    assert lib().orc_is_point_visible(C.byref(flat.desc), _p(pairs), C.c_int(pairs.shape[0]), _p(out)) == 0
    return out.astype(bool)


def camera_rays(params) -> np.ndarray:
    spp = max(1, params.samples_per_side) ** 2
    out = np.zeros((params.width * params.height * spp, 8), dtype=np.float64)
    aa = np.array([params.aa_state, params.aa_inc], dtype=np.uint64)
    assert lib().orc_camera_rays(C.byref(params), _p(aa), _p(out)) == 0
    return out


def pcg_seed(init_state: int, init_seq: int):
    st = np.zeros(2, dtype=np.uint64)
    assert lib().orc_pcg_seed(C.c_uint64(init_state), C.c_uint64(init_seq), _p(st)) == 0
    return int(st[0]), int(st[1])


def pcg_draw(state: int, inc: int, n: int):
    st = np.array([state, inc], dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint32)
    assert lib().orc_pcg_draw(_p(st), C.c_int(n), _p(out)) == 0
    return out, int(st[0])


def pigment_color(flat, pigment: int, uv: np.ndarray) -> np.ndarray:
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((uv.shape[0], 3), dtype=np.float64)
    assert lib().orc_pigment_color(C.byref(flat.desc), C.c_int(pigment), _p(uv), C.c_int(uv.shape[0]), _p(out)) == 0
    return out


def scatter(flat, material: int, inputs: np.ndarray, state: int, inc: int):
    inputs = np.ascontiguousarray(inputs, dtype=np.float64).reshape(-1, 9)
    st = np.array([state, inc], dtype=np.uint64)
    out = np.zeros((inputs.shape[0], 8), dtype=np.float64)
    assert lib().orc_scatter(C.byref(flat.desc), C.c_int(material), _p(inputs), C.c_int(inputs.shape[0]), _p(st), _p(out)) == 0
    return out, int(st[0])


def onb(normals: np.ndarray) -> np.ndarray:
    normals = np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((normals.shape[0], 9), dtype=np.float64)
    assert lib().orc_onb(_p(normals), C.c_int(normals.shape[0]), _p(out)) == 0
    return out
