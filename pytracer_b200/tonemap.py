"""Tone mapping + LDR output on the device: HdrImage.average_luminosity,
normalize_image, clamp_image and write_ldr_image's per-pixel map (hdrimages.py:120-171 of the
reference) as two HBM-bound kernels behind rt_average_luminosity / rt_tone_map (include/rt_api.h,
csrc/rt_tonemap.cu).  No CPU path: without the library or a device these calls raise.

The functions take the fp32 ``(H, W, 3)`` array of an image (``HdrImage.rgb_array()``); the methods of
the same names on :class:`pytracer_b200.hdrimage.HdrImage` forward here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi, _native


def _as_f32(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    if rgb.ndim < 2 or rgb.shape[-1] != 3:
        raise ValueError("expected an array of RGB triples")
    return rgb


def average_luminosity(rgb: np.ndarray, delta: float = 1e-10) -> float:
    """HdrImage.average_luminosity(delta), hdrimages.py:120-128."""
    rgb = _as_f32(rgb)
    lib = _native.require_device()
    out = C.c_double(0.0)
    _native.check(lib.rt_average_luminosity(rgb.ctypes.data, rgb.size // 3, float(delta), 0, None, C.byref(out)))
    return out.value


NORMALIZE, CLAMP = 1, 2  # RT_TONE_* of include/rt_api.h


def tone_map(rgb: np.ndarray, factor: float = 1.0, luminosity=None, gamma: float = 1.0, want_hdr: bool = True,
             want_ldr: bool = True, flags: int = NORMALIZE | CLAMP):
    """normalize_image(factor, luminosity) + clamp_image() [+ the 8-bit map of write_ldr_image].

    Returns ``(hdr float32 array or None, ldr uint8 array or None, stats dict)``; ``luminosity=None``
    (or 0, the reference's ``if not luminosity``) uses the image's own average."""
    rgb = _as_f32(rgb)
    lib = _native.require_device()
    hdr = np.empty_like(rgb) if want_hdr else None
    ldr = np.empty(rgb.shape, dtype=np.uint8) if want_ldr else None
    stats = _abi.rt_tonemap_stats()
    _native.check(lib.rt_tone_map(rgb.ctypes.data, rgb.size // 3, int(flags), float(factor), float(luminosity or 0.0), float(gamma), 0, None,
                                  hdr.ctypes.data if want_hdr else None, ldr.ctypes.data if want_ldr else None, C.byref(stats)))
    return hdr, ldr, stats.as_dict()


def tone_map_device(d_rgb: int, n_pixels: int, factor: float = 1.0, luminosity=None, gamma: float = 1.0, d_out_hdr: int = 0,
                    d_out_ldr: int = 0, stream: int = 0, flags: int = NORMALIZE | CLAMP) -> dict:
    """Same on DEVICE buffers (raw pointers, e.g. ``tensor.data_ptr()``) on a caller stream."""
    lib = _native.require_device()
    stats = _abi.rt_tonemap_stats()
    _native.check(lib.rt_tone_map(d_rgb, int(n_pixels), int(flags), float(factor), float(luminosity or 0.0), float(gamma), 1, stream or None,
                                  d_out_hdr or None, d_out_ldr or None, C.byref(stats)))
    return stats.as_dict()


def write_ldr_image(image, stream, format: str = "PNG", factor: float = 1.0, gamma: float = 1.0, luminosity=None,
                    flags: int = NORMALIZE | CLAMP) -> None:
    """main.py:209-215 in one call: normalise, clamp, quantise on the device, encode with Pillow.
    ``flags=0`` is HdrImage.write_ldr_image alone (hdrimages.py:149-171) on an already tone-mapped image."""
    from PIL import Image

    _, ldr, _ = tone_map(image.rgb_array(), factor, luminosity, gamma, want_hdr=False, flags=flags)
    Image.fromarray(ldr, "RGB").save(stream, format=format)
