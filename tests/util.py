"""Shared helpers of the test-suite: golden fixtures (outputs of the unmodified reference, see
tests/golden/make_golden.py) and scene rebuilding from them."""
import os

import numpy as np

from pytracer_b200 import _abi
from pytracer_b200.flatten import FlatScene
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def camera_from(vec) -> _abi.rt_camera:
    cam = _abi.rt_camera()
    cam.kind, cam.screen_distance, cam.aspect_ratio = int(vec[0]), float(vec[1]), float(vec[2])
    cam.m[:] = [float(v) for v in vec[3:15]]
    return cam


def demo_flat():
    z = golden("demo_scene.npz")
    return FlatScene.from_npz_dict(z), camera_from(z["camera"])


def scene2_flat():
    z = golden("scene2.npz")
    return FlatScene.from_npz_dict(z), camera_from(z["camera"]), camera_from(z["camera_ortho"])


def c1_params(cam, **kw):
    """BASELINE config 1: the reference CLI defaults, 160x120, 1 spp."""
    args = dict(algorithm="pathtracing", samples_per_side=1, num_of_rays=10, max_depth=3, rr_limit=3,
                aa_pcg=PCG(42, 54), pt_pcg=PCG(45, 54))
    args.update(kw)
    return make_params(160, 120, cam, **args)


def luminosity(rgb):
    """Color.luminosity, colors.py:59-61, per pixel."""
    return (rgb.max(axis=-1) + rgb.min(axis=-1)) / 2
