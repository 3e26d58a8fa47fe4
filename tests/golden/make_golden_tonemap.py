#!/usr/bin/env python
"""Golden fixture of the tone-mapping step, made by the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference/src python tests/golden/make_golden_tonemap.py

Input images are fp32 (what HdrImage.write_pfm stores and read_pfm_image gives back): the config-1
path-traced frame and the second scene's point-light frame of the existing fixtures, plus a small
synthetic image spanning 12 decades with exact zeros.  For every case the reference's own
HdrImage.average_luminosity / normalize_image / clamp_image / write_ldr_image run; the PNG bytes are
decoded again with Pillow, so `ldr` is exactly what the reference's `render` / `pfm2png` put on disk.
"""
import io
import os

import numpy as np
from PIL import Image

import pytracer
from pytracer.colors import Color
from pytracer.hdrimages import HdrImage

assert pytracer.__file__.startswith("/root/reference/"), pytracer.__file__
HERE = os.path.dirname(os.path.abspath(__file__))


def to_ref(rgb32):
    h, w, _ = rgb32.shape
    img = HdrImage(w, h)
    flat = rgb32.reshape(-1, 3)
    for i in range(w * h):
        img.pixels[i] = Color(float(flat[i, 0]), float(flat[i, 1]), float(flat[i, 2]))
    return img


def run_case(rgb32, factor, luminosity, gamma):
    img = to_ref(rgb32)
    avg = img.average_luminosity()
    img.normalize_image(factor=factor, luminosity=luminosity)
    img.clamp_image()
    hdr = np.array([(p.r, p.g, p.b) for p in img.pixels], dtype=np.float64).reshape(rgb32.shape)
    buf = io.BytesIO()
    img.write_ldr_image(buf, "PNG", gamma=gamma)
    buf.seek(0)
    ldr = np.array(Image.open(buf).convert("RGB"), dtype=np.uint8)
    return avg, hdr, ldr


def main():
    c1 = np.load(os.path.join(HERE, "demo_c1_pathtracing_160x120.npz"))["rgb"].astype(np.float32)
    s2 = np.load(os.path.join(HERE, "scene2.npz"))["persp_pointlight_rgb"].astype(np.float32)
    rng = np.random.default_rng(20261018)
    syn = (10.0 ** rng.uniform(-8, 4, size=(37, 53, 3))).astype(np.float32)  # odd sizes: ragged tails
    syn[0, :5] = 0.0
    syn[5, 7] = (0.0, 1.0, 0.0)
    images = {"c1": c1, "s2": s2, "syn": syn}
    cases = [  # (image, factor, luminosity (None = the image's own average), gamma)
        ("c1", 1.0, None, 1.0),     # `render` / `demo`: main.py:209-215
        ("c1", 0.7, None, 1.0),     # pfm2png defaults: main.py:218-220
        ("c1", 0.7, 0.5, 2.2),
        ("s2", 1.0, None, 1.0),
        ("s2", 0.18, None, 1.8),
        ("syn", 1.0, None, 1.0),
        ("syn", 2.5, 3.0, 2.2),
    ]
    out = {f"img_{k}": v for k, v in images.items()}
    out["n_cases"] = np.int64(len(cases))
    for i, (name, factor, lum, gamma) in enumerate(cases):
        avg, hdr, ldr = run_case(images[name], factor, lum, gamma)
        out[f"case{i}_image"] = np.array(name)
        out[f"case{i}_params"] = np.array([factor, np.nan if lum is None else lum, gamma], dtype=np.float64)
        out[f"case{i}_avg"] = np.float64(avg)
        out[f"case{i}_hdr"] = hdr
        out[f"case{i}_ldr"] = ldr
        print(name, factor, lum, gamma, "avg luminosity", avg, "ldr mean", ldr.mean())
    # the reference's own known answers (tests/test_all.py:239-268)
    kat = HdrImage(2, 1)
    kat.set_pixel(0, 0, Color(0.5e1, 1.0e1, 1.5e1))
    kat.set_pixel(1, 0, Color(0.5e3, 1.0e3, 1.5e3))
    out["kat_avg_delta0"] = np.float64(kat.average_luminosity(delta=0.0))
    path = os.path.join(HERE, "tonemap.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
