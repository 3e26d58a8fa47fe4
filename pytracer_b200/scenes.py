"""Benchmark scenes built programmatically (the GPU box has no copy of the reference tree).

* :func:`demo_scene` — the content of the reference's ``examples/demo.txt`` (28 lines: two planes,
  one mirror sphere, one point light, perspective camera), constructed with the same chain of
  ``Transformation`` products ``parse_transformation`` performs (scene_file.py:517-565), so the
  flattened matrices carry the same bits as a scene parsed by the reference
  (pinned by tests/golden/demo_scene.npz).
* :func:`random_spheres_scene` — BASELINE configs 4/5: N randomly transformed
  ellipsoids + ground and sky planes, seeded by the reference's own PCG so that any host
  reproduces the same scene.  ``to_scene_text`` writes the same scene in the reference's scene
  language (fixed-point literals: its lexer cannot read negative exponents).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .hdrimage import HdrImage
from .pcg import PCG
from .scene import (
    BLACK, CheckeredPigment, Color, DiffuseBRDF, ImagePigment, Material, PerspectiveCamera, Plane,
    Point, PointLight, SpecularBRDF, Sphere, Transformation, UniformPigment, Vec, World,
    rotation_x, rotation_y, rotation_z, scaling, translation,
)


def demo_scene(clock: float = 150.0) -> Tuple[World, PerspectiveCamera]:
    sky = Material(DiffuseBRDF(UniformPigment(Color(0, 0, 0))), UniformPigment(Color(0.7, 0.5, 1)))
    ground = Material(
        DiffuseBRDF(CheckeredPigment(Color(0.3, 0.5, 0.1), Color(0.1, 0.2, 0.5), 4)),
        UniformPigment(Color(0, 0, 0)),
    )
    mirror = Material(SpecularBRDF(UniformPigment(Color(0.5, 0.5, 0.5))), UniformPigment(Color(0, 0, 0)))
    world = World()
    world.add_light(PointLight(Point(10, 10, 10), Color(1, 1, 1), 1))
    ident = Transformation()
    world.add_shape(Plane(ident * translation(Vec(0, 0, 100)) * rotation_y(clock), sky))
    world.add_shape(Plane(Transformation(), ground))
    world.add_shape(Sphere(ident * translation(Vec(0, 0, 1)), mirror))
    camera = PerspectiveCamera(
        screen_distance=1.0, aspect_ratio=1.0,
        transformation=ident * rotation_z(30) * translation(Vec(-4, 0, 1)),
    )
    return world, camera


def _r6(x: float) -> float:
    """Round through the '%.6f' literal the scene text carries, so text and objects agree."""
    return float(f"{x:.6f}")


def gradient_texture(width: int = 512, height: int = 256) -> HdrImage:
    """texel(x, y) = (x/(w-1), y/(h-1), 0.5), stored as fp32 like a PFM would."""
    xs = (np.arange(width, dtype=np.float64) / (width - 1)).astype(np.float32)
    ys = (np.arange(height, dtype=np.float64) / (height - 1)).astype(np.float32)
    rgb = np.empty((height, width, 3), dtype=np.float32)
    rgb[..., 0] = xs[None, :]
    rgb[..., 1] = ys[:, None]
    rgb[..., 2] = 0.5
    return HdrImage.from_array(rgb)


class RandomSpheres:
    """Result of :func:`random_spheres_scene`."""

    def __init__(self, world, camera, sphere_params, palette_names, texture, extent, with_light):
        self.world, self.camera = world, camera
        self.sphere_params, self.palette_names = sphere_params, palette_names
        self.texture, self.extent, self.with_light = texture, extent, with_light

    def to_scene_text(self, texture_path: str = "texture.pfm") -> str:
        lines: List[str] = [
            "material m0(diffuse(uniform(<0.800000, 0.300000, 0.300000>)), uniform(<0, 0, 0>))",
            "material m1(diffuse(uniform(<0.300000, 0.800000, 0.300000>)), uniform(<0, 0, 0>))",
            "material m2(diffuse(uniform(<0.300000, 0.300000, 0.800000>)), uniform(<0, 0, 0>))",
            "material m3(diffuse(checkered(<0.900000, 0.900000, 0.100000>, <0.100000, 0.100000, 0.900000>, 8)), uniform(<0, 0, 0>))",
            "material m4(diffuse(checkered(<0.900000, 0.100000, 0.900000>, <0.100000, 0.900000, 0.900000>, 8)), uniform(<0, 0, 0>))",
            f'material m5(diffuse(image("{texture_path}")), uniform(<0, 0, 0>))',
            "material m6(specular(uniform(<0.600000, 0.600000, 0.600000>)), uniform(<0, 0, 0>))",
            "material m7(specular(uniform(<0.900000, 0.700000, 0.400000>)), uniform(<0, 0, 0>))",
            "material ground(diffuse(checkered(<0.300000, 0.500000, 0.100000>, <0.100000, 0.200000, 0.500000>, 2)), uniform(<0, 0, 0>))",
            "material sky(diffuse(uniform(<0, 0, 0>)), uniform(<1, 1, 1>))",
        ]
        if self.with_light:
            lines.append("point_light([30, 30, 60], <1, 1, 1>, 0)")
        for (x, y, z, a, b, c, sx, sy, sz), name in zip(self.sphere_params, self.palette_names):
            lines.append(
                f"sphere({name}, translation([{x:.6f}, {y:.6f}, {z:.6f}]) * rotation_z({a:.6f}) * "
                f"rotation_y({b:.6f}) * rotation_x({c:.6f}) * scaling([{sx:.6f}, {sy:.6f}, {sz:.6f}]))"
            )
        lines.append("plane(ground, identity)")
        lines.append("plane(sky, translation([0, 0, 100]))")
        lines.append("camera(perspective, translation([-25, 0, 6]) * rotation_y(12), 1.777778, 1.0)")
        return "\n".join(lines) + "\n"


def random_spheres_scene(n_spheres: int = 1024, init_state: int = 2024, init_seq: int = 4,
                         extent: float = 20.0, with_light: bool = False) -> RandomSpheres:
    """``random_spheres_scene(1024, 2024, 4, 20.0)``;
    C5: ``random_spheres_scene(4096, 2025, 5, 40.0, with_light=True)``."""
    pcg = PCG(init_state, init_seq)

    def uniform(lo: float, hi: float) -> float:
        return _r6(lo + (hi - lo) * pcg.random_float())

    texture = gradient_texture()
    palette = [
        Material(DiffuseBRDF(UniformPigment(Color(0.8, 0.3, 0.3))), UniformPigment(BLACK)),
        Material(DiffuseBRDF(UniformPigment(Color(0.3, 0.8, 0.3))), UniformPigment(BLACK)),
        Material(DiffuseBRDF(UniformPigment(Color(0.3, 0.3, 0.8))), UniformPigment(BLACK)),
        Material(DiffuseBRDF(CheckeredPigment(Color(0.9, 0.9, 0.1), Color(0.1, 0.1, 0.9), 8)), UniformPigment(BLACK)),
        Material(DiffuseBRDF(CheckeredPigment(Color(0.9, 0.1, 0.9), Color(0.1, 0.9, 0.9), 8)), UniformPigment(BLACK)),
        Material(DiffuseBRDF(ImagePigment(texture)), UniformPigment(BLACK)),
        Material(SpecularBRDF(UniformPigment(Color(0.6, 0.6, 0.6))), UniformPigment(BLACK)),
        Material(SpecularBRDF(UniformPigment(Color(0.9, 0.7, 0.4))), UniformPigment(BLACK)),
    ]
    ground = Material(DiffuseBRDF(CheckeredPigment(Color(0.3, 0.5, 0.1), Color(0.1, 0.2, 0.5), 2)), UniformPigment(BLACK))
    sky = Material(DiffuseBRDF(UniformPigment(Color(0, 0, 0))), UniformPigment(Color(1, 1, 1)))

    world = World()
    if with_light:
        world.add_light(PointLight(Point(30, 30, 60), Color(1, 1, 1), 0))
    params, names = [], []
    for i in range(n_spheres):
        x, y = uniform(-extent, extent), uniform(-extent, extent)
        z = uniform(0.2, 6.0)
        a, b, c = uniform(0.0, 360.0), uniform(0.0, 360.0), uniform(0.0, 360.0)
        sx, sy, sz = uniform(0.15, 0.6), uniform(0.15, 0.6), uniform(0.15, 0.6)
        t = (Transformation() * translation(Vec(x, y, z)) * rotation_z(a) * rotation_y(b)
             * rotation_x(c) * scaling(Vec(sx, sy, sz)))
        world.add_shape(Sphere(t, palette[i % len(palette)]))
        params.append((x, y, z, a, b, c, sx, sy, sz))
        names.append(f"m{i % len(palette)}")
    world.add_shape(Plane(Transformation(), ground))
    world.add_shape(Plane(Transformation() * translation(Vec(0, 0, 100)), sky))
    camera = PerspectiveCamera(
        screen_distance=1.0, aspect_ratio=1.777778,
        transformation=Transformation() * translation(Vec(-25, 0, 6)) * rotation_y(12),
    )
    return RandomSpheres(world, camera, params, names, texture, extent, with_light)
