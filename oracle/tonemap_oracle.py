"""TEST INFRASTRUCTURE — CPU restatement of the reference's tone mapping (numpy, fp64).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this
module; the product (pytracer_b200/tonemap.py) runs the CUDA kernels of csrc/rt_tonemap.cu and has no
CPU path.  Pinned: bit-exact (LDR bytes, average luminosity) and <= 1e-15 relative (normalised values) against
outputs of the unmodified reference, tests/golden/tonemap.npz (made by
tests/golden/make_golden_tonemap.py); see tests/test_oracle_golden.py.

Follows /root/reference/src/pytracer/hdrimages.py:
    average_luminosity :120-128   10 ** (sum(log10(delta + luminosity)) / len(pixels))
    normalize_image    :130-140   pixel * (factor / luminosity);  `if not luminosity` -> average
    clamp_image        :142-147   x / (1 + x)   (_clamp, :50-52)
    write_ldr_image    :149-171   int(255 * pow(c, 1 / gamma)) per channel
and colors.py:59-61 (luminosity = (max + min) / 2).
"""
from __future__ import annotations

import math

import numpy as np


def luminosity(rgb: np.ndarray) -> np.ndarray:
    rgb = np.asarray(rgb, dtype=np.float64)
    return (rgb.max(axis=-1) + rgb.min(axis=-1)) / 2


def average_luminosity(rgb: np.ndarray, delta: float = 1e-10) -> float:
    lum = luminosity(rgb).reshape(-1)
    # math.log10 per pixel and a sequential fp64 running sum, exactly the reference's loop
    # (np.cumsum adds in order; numpy's own log10 may differ from libm's in the last bit)
    logs = np.array([math.log10(v) for v in (delta + lum).tolist()], dtype=np.float64)
    cumsum = float(np.cumsum(logs)[-1])
    return math.pow(10, cumsum / lum.size)


def normalize_and_clamp(rgb: np.ndarray, factor: float = 1.0, luminosity_value=None, delta: float = 1e-10) -> np.ndarray:
    rgb = np.asarray(rgb, dtype=np.float64)
    lum = luminosity_value if luminosity_value else average_luminosity(rgb, delta)
    x = rgb * (factor / lum)
    return x / (1 + x)


def ldr_bytes(clamped: np.ndarray, gamma: float = 1.0) -> np.ndarray:
    g = clamped if gamma == 1.0 else np.power(clamped, 1 / gamma)
    return np.clip((255 * g).astype(np.int64), 0, 255).astype(np.uint8)


def tone_map(rgb: np.ndarray, factor: float = 1.0, luminosity_value=None, gamma: float = 1.0):
    """(average luminosity used, normalised + clamped fp64 image, uint8 LDR image)."""
    rgb = np.asarray(rgb, dtype=np.float64)
    lum = luminosity_value if luminosity_value else average_luminosity(rgb)
    hdr = normalize_and_clamp(rgb, factor, lum)
    return lum, hdr, ldr_bytes(hdr, gamma)
