"""Tone mapping + LDR output for the CLI (hdrimages.py:120-171 of the reference: log-average
luminosity, normalisation by factor/luminosity, x/(1+x) clamp, gamma, 8-bit PNG), vectorised with
numpy on the host.  SURVEY §8(f)-2 lists a device version as a later row; this is the host-side
stand-in so that `render` produces the same PNG as the reference without its per-pixel Python loops."""
from __future__ import annotations

import numpy as np


def average_luminosity(rgb: np.ndarray, delta: float = 1e-10) -> float:
    lum = (rgb.max(axis=-1) + rgb.min(axis=-1)) / 2  # Color.luminosity, colors.py:59-61
    return float(10 ** np.mean(np.log10(delta + lum.astype(np.float64))))


def tone_map(rgb: np.ndarray, factor: float = 1.0, luminosity=None) -> np.ndarray:
    rgb = rgb.astype(np.float64)
    lum = luminosity if luminosity else average_luminosity(rgb)
    rgb = rgb * (factor / lum)
    return rgb / (1 + rgb)


def write_ldr_image(image, stream, format: str = "PNG", factor: float = 1.0, gamma: float = 1.0, luminosity=None) -> None:
    from PIL import Image

    ldr = tone_map(image.rgb_array(), factor, luminosity)
    data = (255 * np.power(ldr, 1 / gamma)).astype(np.int64).clip(0, 255).astype(np.uint8)
    Image.fromarray(data, "RGB").save(stream, format=format)
