"""World / Camera object graphs -> the flat buffers of include/rt_api.h.

Duck-typed on class names (looked up along the MRO) and attributes, so it takes the reference's
own objects (``pytracer.world.World`` filled by ``pytracer.scene_file.parse_scene``) as well as
the records of :mod:`pytracer_b200.scene`.  What it reads, with the reference definitions:

* ``world.shapes`` (world.py:39) in order — ``Sphere`` / ``Plane`` (shapes.py:88,154),
  ``shape.transformation.m`` / ``.invm`` (transformations.py:54-56; rows 0..2 of the 4x4 lists),
  ``shape.material`` (materials.py:199-204), de-duplicated by object identity;
* ``material.brdf`` — ``DiffuseBRDF`` / ``SpecularBRDF`` with ``.pigment`` (+ ``threshold_angle_rad``),
  ``material.emitted_radiance``;
* pigments: ``UniformPigment.color``, ``CheckeredPigment.color1/.color2/.num_of_steps``,
  ``ImagePigment.image`` (``width``, ``height``, ``pixels`` row-major from the top row);
* ``world.point_lights`` (lights.py:25-39); cameras (camera.py:42-124).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List

import itertools

import numpy as np

from . import _abi


def _kind_of(obj, table: Dict[str, int], what: str) -> int:
    for klass in type(obj).__mro__:
        if klass.__name__ in table:
            return table[klass.__name__]
    raise TypeError(f"unsupported {what}: {type(obj).__name__} (supported: {', '.join(table)})")


_SHAPES = {"Sphere": _abi.RT_SHAPE_SPHERE, "Plane": _abi.RT_SHAPE_PLANE}
_BRDFS = {"DiffuseBRDF": _abi.RT_BRDF_DIFFUSE, "SpecularBRDF": _abi.RT_BRDF_SPECULAR}
_PIGMENTS = {
    "UniformPigment": _abi.RT_PIGMENT_UNIFORM,
    "CheckeredPigment": _abi.RT_PIGMENT_CHECKERED,
    "ImagePigment": _abi.RT_PIGMENT_IMAGE,
}
_CAMERAS = {"OrthogonalCamera": _abi.RT_CAMERA_ORTHOGONAL, "PerspectiveCamera": _abi.RT_CAMERA_PERSPECTIVE}


def _rows3x4(matrix) -> List[float]:
    return [float(matrix[i][j]) for i in range(3) for j in range(4)]


def _rgb(color) -> tuple:
    return (float(color.r), float(color.g), float(color.b))


def image_to_array(image) -> np.ndarray:
    """(height, width, 3) float64, row 0 = top, from any HdrImage-like object."""
    arr = getattr(image, "rgb_array", None)
    if arr is not None:
        return np.asarray(arr() if callable(arr) else arr, dtype=np.float64).reshape(image.height, image.width, 3)
    flat = np.fromiter((c for p in image.pixels for c in (p.r, p.g, p.b)), dtype=np.float64,
                       count=3 * image.width * image.height)
    return flat.reshape(image.height, image.width, 3)


@dataclass
class FlatScene:
    """SoA host buffers + the rt_scene_desc pointing into them (keep this object alive while the
    descriptor is in use)."""

    shape_kind: np.ndarray
    shape_material: np.ndarray
    shape_m: np.ndarray
    shape_invm: np.ndarray
    materials: C.Array
    pigments: C.Array
    lights: C.Array
    texels: np.ndarray
    desc: _abi.rt_scene_desc = field(default=None)

    @property
    def n_shapes(self) -> int:
        return int(self.shape_kind.shape[0])

    def build_desc(self) -> _abi.rt_scene_desc:
        d = _abi.rt_scene_desc()
        d.n_shapes = self.n_shapes
        d.n_materials = len(self.materials)
        d.n_pigments = len(self.pigments)
        d.n_lights = len(self.lights)
        d.shape_kind = self.shape_kind.ctypes.data
        d.shape_material = self.shape_material.ctypes.data
        d.shape_m = self.shape_m.ctypes.data
        d.shape_invm = self.shape_invm.ctypes.data
        d.materials = C.addressof(self.materials) if len(self.materials) else None
        d.pigments = C.addressof(self.pigments) if len(self.pigments) else None
        d.lights = C.addressof(self.lights) if len(self.lights) else None
        d.n_texels = int(self.texels.shape[0])
        d.texels = self.texels.ctypes.data if self.texels.size else None
        self.desc = d
        return d

    def differs_only_in_transforms(self, other: "FlatScene") -> bool:
        """True when `other` has the same shapes, materials, pigments, lights and textures and may differ
        in Transformation.m / .invm only (the next frame of an animation)."""
        return (np.array_equal(self.shape_kind, other.shape_kind) and np.array_equal(self.shape_material, other.shape_material)
                and bytes(self.materials) == bytes(other.materials) and bytes(self.pigments) == bytes(other.pigments)
                and bytes(self.lights) == bytes(other.lights) and self.texels.shape == other.texels.shape
                and np.array_equal(self.texels, other.texels))

    def to_npz_dict(self) -> dict:
        """Portable dump (used for the golden fixtures)."""
        mats = np.array([(m.brdf_kind, m.brdf_pigment, m.emitted_pigment) for m in self.materials], dtype=np.int32).reshape(-1, 3)
        thr = np.array([m.threshold_angle_rad for m in self.materials], dtype=np.float64)
        pig_i = np.array([(p.kind, p.num_of_steps, p.tex_width, p.tex_height, p.tex_offset) for p in self.pigments], dtype=np.int64).reshape(-1, 5)
        pig_c = np.array([list(p.color1) + list(p.color2) for p in self.pigments], dtype=np.float64).reshape(-1, 6)
        lights = np.array([list(l.position) + list(l.color) + [l.linear_radius] for l in self.lights], dtype=np.float64).reshape(-1, 7)
        return dict(shape_kind=self.shape_kind, shape_material=self.shape_material, shape_m=self.shape_m,
                    shape_invm=self.shape_invm, materials=mats, thresholds=thr, pigments_i=pig_i,
                    pigments_c=pig_c, lights=lights, texels=self.texels)

    @staticmethod
    def from_npz_dict(z) -> "FlatScene":
        mats = (_abi.rt_material * len(z["materials"]))()
        for m, (bk, bp, ep), thr in zip(mats, z["materials"], z["thresholds"]):
            m.brdf_kind, m.brdf_pigment, m.emitted_pigment, m.threshold_angle_rad = int(bk), int(bp), int(ep), float(thr)
        pigs = (_abi.rt_pigment * len(z["pigments_i"]))()
        for p, ints, cols in zip(pigs, z["pigments_i"], z["pigments_c"]):
            p.kind, p.num_of_steps, p.tex_width, p.tex_height, p.tex_offset = (int(v) for v in ints)
            p.color1[:] = [float(v) for v in cols[:3]]
            p.color2[:] = [float(v) for v in cols[3:]]
        lights = (_abi.rt_light * len(z["lights"]))()
        for l, row in zip(lights, z["lights"]):
            l.position[:] = [float(v) for v in row[0:3]]
            l.color[:] = [float(v) for v in row[3:6]]
            l.linear_radius = float(row[6])
        fs = FlatScene(
            shape_kind=np.ascontiguousarray(z["shape_kind"], dtype=np.int32),
            shape_material=np.ascontiguousarray(z["shape_material"], dtype=np.int32),
            shape_m=np.ascontiguousarray(z["shape_m"], dtype=np.float64),
            shape_invm=np.ascontiguousarray(z["shape_invm"], dtype=np.float64),
            materials=mats, pigments=pigs, lights=lights,
            texels=np.ascontiguousarray(z["texels"], dtype=np.float64).reshape(-1, 3),
        )
        fs.build_desc()
        return fs


def flatten_world(world) -> FlatScene:
    pigment_index: Dict[int, int] = {}
    pigment_recs: List[dict] = []
    texel_blocks: List[np.ndarray] = []
    n_texels = 0
    keepalive = []  # ids are only unique while the objects live

    def add_pigment(pig) -> int:
        nonlocal n_texels
        key = id(pig)
        if key in pigment_index:
            return pigment_index[key]
        kind = _kind_of(pig, _PIGMENTS, "pigment")
        rec = dict(kind=kind, steps=0, w=0, h=0, off=0, c1=(0.0, 0.0, 0.0), c2=(0.0, 0.0, 0.0))
        if kind == _abi.RT_PIGMENT_UNIFORM:
            rec["c1"] = _rgb(pig.color)
        elif kind == _abi.RT_PIGMENT_CHECKERED:
            rec["c1"], rec["c2"], rec["steps"] = _rgb(pig.color1), _rgb(pig.color2), int(pig.num_of_steps)
        else:
            arr = image_to_array(pig.image)
            rec["w"], rec["h"], rec["off"] = int(pig.image.width), int(pig.image.height), n_texels
            texel_blocks.append(arr.reshape(-1, 3))
            n_texels += arr.shape[0] * arr.shape[1]
        pigment_index[key] = len(pigment_recs)
        pigment_recs.append(rec)
        keepalive.append(pig)
        return pigment_index[key]

    material_index: Dict[int, int] = {}
    material_recs: List[tuple] = []

    def add_material(mat) -> int:
        key = id(mat)
        if key in material_index:
            return material_index[key]
        brdf = mat.brdf
        kind = _kind_of(brdf, _BRDFS, "BRDF")
        thr = float(getattr(brdf, "threshold_angle_rad", 0.0))
        material_recs.append((kind, add_pigment(brdf.pigment), add_pigment(mat.emitted_radiance), thr))
        material_index[key] = len(material_recs) - 1
        keepalive.append(mat)
        return material_index[key]

    shapes = list(world.shapes)
    n = len(shapes)
    kind_of_type: Dict[type, int] = {}
    kinds, smats, ms, invms = [], [], [], []
    for shape in shapes:  # one pass of attribute reads; the matrices are converted in bulk below
        t = type(shape)
        k = kind_of_type.get(t)
        if k is None:
            k = kind_of_type[t] = _kind_of(shape, _SHAPES, "shape")
        kinds.append(k)
        smats.append(add_material(shape.material))
        tr = shape.transformation
        ms.append(tr.m)
        invms.append(tr.invm)
    kind = np.array(kinds, dtype=np.int32).reshape(n)
    smat = np.array(smats, dtype=np.int32).reshape(n)

    def rows(mats) -> np.ndarray:  # n x (4x4 nested lists) -> (n, 12): rows 0..2
        if not mats:
            return np.zeros((0, 12), dtype=np.float64)
        # (np.fromiter over the chained rows is 2.4x faster than np.array on the nested lists: 4 ms instead of
        # 10 ms for the 4 098 shapes of config 5, twice per flatten, and a renderer flattens its World per image)
        chain = itertools.chain.from_iterable
        try:
            flat = np.fromiter(chain(chain(mats)), dtype=np.float64, count=n * 16)
        except (TypeError, ValueError):  # not 4x4 nested sequences of numbers: let numpy say what is wrong
            flat = np.array(mats, dtype=np.float64).reshape(-1)
        return np.ascontiguousarray(flat.reshape(n, 4, 4)[:, :3, :].reshape(n, 12))

    m, invm = rows(ms), rows(invms)

    mats = (_abi.rt_material * len(material_recs))()
    for dst, (bk, bp, ep, thr) in zip(mats, material_recs):
        dst.brdf_kind, dst.brdf_pigment, dst.emitted_pigment, dst.threshold_angle_rad = bk, bp, ep, thr
    pigs = (_abi.rt_pigment * len(pigment_recs))()
    for dst, rec in zip(pigs, pigment_recs):
        dst.kind, dst.num_of_steps = rec["kind"], rec["steps"]
        dst.tex_width, dst.tex_height, dst.tex_offset = rec["w"], rec["h"], rec["off"]
        dst.color1[:] = rec["c1"]
        dst.color2[:] = rec["c2"]
    point_lights = list(getattr(world, "point_lights", []))
    lights = (_abi.rt_light * len(point_lights))()
    for dst, light in zip(lights, point_lights):
        dst.position[:] = (float(light.position.x), float(light.position.y), float(light.position.z))
        dst.color[:] = _rgb(light.color)
        dst.linear_radius = float(light.linear_radius)

    texels = np.concatenate(texel_blocks, axis=0) if texel_blocks else np.zeros((0, 3), dtype=np.float64)
    fs = FlatScene(kind, smat, m, invm, mats, pigs, lights, np.ascontiguousarray(texels, dtype=np.float64))
    fs.build_desc()
    return fs


def flatten_camera(camera) -> _abi.rt_camera:
    cam = _abi.rt_camera()
    cam.kind = _kind_of(camera, _CAMERAS, "camera")
    cam.screen_distance = float(getattr(camera, "screen_distance", 1.0))
    cam.aspect_ratio = float(camera.aspect_ratio)
    cam.m[:] = _rows3x4(camera.transformation.m)
    return cam
