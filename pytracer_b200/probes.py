"""Single-item evaluations behind the methods of :mod:`pytracer_b200.scene`.

Each function flattens the objects involved, launches the corresponding device probe (the same
device functions the render kernels call, fp64 instantiation) and wraps the answer in the
reference's result types.  They exist so that the reference's unit tests can be restated against
the CUDA code one function at a time; a renderer never goes through them.
"""
from __future__ import annotations

import math

import numpy as np

from . import device
from .flatten import flatten_camera
from .scene import (Color, HitRecord, Material, Normal, Point, Ray, Vec, Vec2d, World, UniformPigment,
                    DiffuseBRDF)


def _ray8(ray: Ray) -> np.ndarray:
    return np.array([[ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x, ray.dir.y, ray.dir.z, ray.tmin, ray.tmax]])


def _ray_from8(row, depth: int = 0) -> Ray:
    return Ray(Point(*row[0:3]), Vec(*row[3:6]), float(row[6]), float(row[7]), depth)


def world_ray_intersection(world, ray: Ray, normalize: bool = True, precision: str = "f64"):
    if not world.shapes:
        return None
    scene = device.DeviceScene(world)
    try:
        hit = scene.intersect(_ray8(ray), precision, normalize)[0]
    finally:
        scene.close()
    if hit.shape < 0:
        return None
    return HitRecord(
        world_point=Point(*hit.world_point), normal=Normal(*hit.normal), surface_point=Vec2d(*hit.uv),
        t=hit.t, ray=ray, material=world.shapes[hit.shape].material,
    )


def quick_ray_intersection(world, ray: Ray, precision: str = "f64") -> bool:
    # Shape.quick_ray_intersection(ray) is the any-hit test on the segment (tmin, tmax) of `ray`;
    # World.is_point_visible builds Ray(P, L-P, 1e-2/|L-P|, 1) and negates it.  The probe takes the
    # same form, so feed it the closest-hit kernel with the ray's own limits instead.
    scene = device.DeviceScene(world)
    try:
        hit = scene.intersect(_ray8(ray), precision, False)[0]
    finally:
        scene.close()
    return hit.shape >= 0


def world_is_point_visible(world, point: Point, observer_pos: Point, precision: str = "f64") -> bool:
    if not world.shapes:
        return True
    scene = device.DeviceScene(world)
    try:
        pairs = np.array([[point.x, point.y, point.z, observer_pos.x, observer_pos.y, observer_pos.z]])
        return bool(scene.is_point_visible(pairs, precision)[0])
    finally:
        scene.close()


def _one_material_world(material) -> World:
    from .scene import Sphere

    w = World()
    w.add_shape(Sphere(material=material))
    return w


def pigment_get_color(pigment, uv: Vec2d, precision: str = "f64") -> Color:
    material = Material(DiffuseBRDF(pigment), UniformPigment(Color()))
    scene = device.DeviceScene(_one_material_world(material))
    try:
        rgb = scene.pigment_color(scene.flat.materials[0].brdf_pigment, np.array([[uv.u, uv.v]]), precision)[0]
    finally:
        scene.close()
    return Color(*rgb)


def brdf_scatter_ray(brdf, pcg, incoming_dir: Vec, interaction_point: Point, normal: Normal, depth: int,
                     precision: str = "f64") -> Ray:
    scene = device.DeviceScene(_one_material_world(Material(brdf, UniformPigment(Color()))))
    try:
        row = np.array([[*incoming_dir.xyz(), *interaction_point.xyz(), *normal.xyz()]])
        out, state = scene.scatter(0, row, pcg.state, pcg.inc, precision)
    finally:
        scene.close()
    pcg.state = state
    return _ray_from8(out[0], depth)


def camera_fire_ray(camera, u: float, v: float, precision: str = "f64") -> Ray:
    out = device.camera_fire(flatten_camera(camera), np.array([[u, v]]), precision)
    return _ray_from8(out[0])


def onb_from_z(normal, precision: str = "f64"):
    out = device.onb(np.array([[normal.x, normal.y, normal.z]]), precision)[0]
    return Vec(*out[0:3]), Vec(*out[3:6]), Vec(*out[6:9])
