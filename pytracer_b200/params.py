"""Builds the rt_render_params struct from the arguments the reference spreads over
ImageTracer (imagetracer.py:29-46), the Renderer constructors (render.py:31-155) and the CLI."""
from __future__ import annotations

from typing import Optional

from . import _abi
from .flatten import flatten_camera
from .pcg import PCG


def _set3(dst, color) -> None:
    if hasattr(color, "r"):
        dst[:] = (float(color.r), float(color.g), float(color.b))
    else:
        dst[:] = tuple(float(c) for c in color)


def make_params(
    width: int,
    height: int,
    camera,
    algorithm="pathtracing",
    samples_per_side: int = 0,
    background=(0.0, 0.0, 0.0),
    onoff_color=(1.0, 1.0, 1.0),
    ambient=(0.1, 0.1, 0.1),
    num_of_rays: int = 10,
    max_depth: int = 10,
    rr_limit: int = 3,
    aa_pcg: Optional[PCG] = None,
    pt_pcg: Optional[PCG] = None,
    rng_mode: int = _abi.RT_RNG_STREAMS,
    part_mode: int = _abi.RT_PART_NONE,
    part_rank: int = 0,
    part_count: int = 1,
    variant="auto",
    precision="auto",
    out_f64: bool = False,
    hit_mode: int = _abi.RT_HIT_SHAPE,
    accel="none",
    rows_layout: int = _abi.RT_ROWS_FULL,
) -> _abi.rt_render_params:
    p = _abi.rt_render_params()
    p.width, p.height, p.samples_per_side = int(width), int(height), int(samples_per_side)
    p.algorithm = _abi.ALGORITHMS[algorithm] if isinstance(algorithm, str) else int(algorithm)
    p.camera = camera if isinstance(camera, _abi.rt_camera) else flatten_camera(camera)
    _set3(p.background, background)
    _set3(p.onoff_color, onoff_color)
    _set3(p.ambient, ambient)
    p.num_of_rays, p.max_depth, p.rr_limit = int(num_of_rays), int(max_depth), int(rr_limit)
    p.rng_mode = rng_mode
    aa_pcg = aa_pcg if aa_pcg is not None else PCG()
    pt_pcg = pt_pcg if pt_pcg is not None else PCG()
    p.aa_state, p.aa_inc = aa_pcg.state, aa_pcg.inc
    p.pt_state, p.pt_inc = pt_pcg.state, pt_pcg.inc
    p.replay_states = None
    p.part_mode, p.part_rank, p.part_count = part_mode, part_rank, part_count
    p.variant = _abi.VARIANTS[variant] if isinstance(variant, str) else int(variant)
    p.precision = _abi.PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    p.out_f64 = 1 if out_f64 else 0
    p.hit_mode = hit_mode
    p.accel = _abi.ACCELS[accel] if isinstance(accel, str) else int(accel)
    p.rows_layout = int(rows_layout)
    p.n_peer_images = 0
    return p
