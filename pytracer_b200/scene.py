"""Host-side scene description: the names, constructor arguments and error behaviour of the
reference's scene model, as plain records.

The reference (ziotom78/pytracer) builds a graph of Python objects and walks it for every ray.
Here the same names only *describe* a scene; :func:`pytracer_b200.flatten.flatten_world` turns the
graph into SoA buffers and every bit of per-ray arithmetic runs in the CUDA library.  The methods
that evaluate something (``ray_intersection``, ``get_color``, ``scatter_ray``, ``fire_ray`` …) are
single-item launches of the same device code the renderer uses — there is no CPU fallback.

``flatten_world`` is duck-typed on class *names* and attributes, so objects built by the
reference's own package (``pytracer.world.World`` from ``pytracer.scene_file.parse_scene``) are
accepted unchanged: that is what makes the CUDA renderer a drop-in.

Reference counterparts (``src/pytracer/``): colors.py:22-79, geometry.py:61-245,
transformations.py:47-233, ray.py:28-69, hitrecord.py:27-46, materials.py:36-204,
shapes.py:57-198, lights.py:25-39, camera.py:42-124, world.py:26-80.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence


# ----------------------------------------------------------------------------- colours
@dataclass
class Color:
    """RGB triple (colors.py:22-69)."""

    r: float = 0.0
    g: float = 0.0
    b: float = 0.0

    def __add__(self, other: "Color") -> "Color":
        return Color(self.r + other.r, self.g + other.g, self.b + other.b)

    def __mul__(self, other) -> "Color":
        if hasattr(other, "r"):
            return Color(self.r * other.r, self.g * other.g, self.b * other.b)
        return Color(self.r * other, self.g * other, self.b * other)

    def luminosity(self) -> float:
        return (max(self.r, self.g, self.b) + min(self.r, self.g, self.b)) / 2

    def is_close(self, other: "Color", epsilon: float = 1e-6) -> bool:
        return all(abs(a - b) < epsilon for a, b in zip(self.rgb(), other.rgb()))

    def rgb(self):
        return (self.r, self.g, self.b)


BLACK = Color(0.0, 0.0, 0.0)
WHITE = Color(1.0, 1.0, 1.0)


# ----------------------------------------------------------------------------- geometry
class _XYZ:
    """Shared behaviour of the three xyz triples; the distinction between them only matters to
    Transformation.__mul__ (transformations.py:58-86), which runs on the device here."""

    __slots__ = ("x", "y", "z")

    def __init__(self, x: float = 0.0, y: float = 0.0, z: float = 0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def xyz(self):
        return (self.x, self.y, self.z)

    def __iter__(self):
        return iter(self.xyz())

    def __getitem__(self, i: int) -> float:
        if not 0 <= i < 3:
            raise AssertionError(f"wrong vector index {i}")
        return self.xyz()[i]

    def __eq__(self, other) -> bool:
        return type(other) is type(self) and self.xyz() == other.xyz()

    def __repr__(self) -> str:
        return f"{type(self).__name__}(x={self.x}, y={self.y}, z={self.z})"

    def is_close(self, other, epsilon: float = 1e-5) -> bool:
        assert isinstance(other, type(self))
        return all(abs(a - b) < epsilon for a, b in zip(self.xyz(), other.xyz()))

    def __neg__(self):
        return type(self)(-self.x, -self.y, -self.z)

    def __mul__(self, scalar: float):
        return type(self)(scalar * self.x, scalar * self.y, scalar * self.z)


class Vec(_XYZ):
    def __add__(self, other):
        if isinstance(other, (Vec, Point)):
            return type(other)(self.x + other.x, self.y + other.y, self.z + other.z)
        raise TypeError(f"Unable to run Vec.__add__ on a {type(self)} and a {type(other)}.")

    def __sub__(self, other):
        if isinstance(other, Vec):
            return Vec(self.x - other.x, self.y - other.y, self.z - other.z)
        raise TypeError(f"Unable to run Vec.__sub__ on a {type(self)} and a {type(other)}.")

    def dot(self, other) -> float:
        return self.x * other.x + self.y * other.y + self.z * other.z

    def cross(self, other) -> "Vec":
        return Vec(
            self.y * other.z - self.z * other.y,
            self.z * other.x - self.x * other.z,
            self.x * other.y - self.y * other.x,
        )

    def squared_norm(self) -> float:
        return self.dot(self)

    def norm(self) -> float:
        return math.sqrt(self.squared_norm())

    def normalize(self) -> "Vec":
        n = self.norm()
        self.x, self.y, self.z = self.x / n, self.y / n, self.z / n
        return self


class Point(_XYZ):
    def __add__(self, other):
        if isinstance(other, Vec):
            return Point(self.x + other.x, self.y + other.y, self.z + other.z)
        raise TypeError(f"Unable to run Point.__add__ on a {type(self)} and a {type(other)}.")

    def __sub__(self, other):
        if isinstance(other, Vec):
            return Point(self.x - other.x, self.y - other.y, self.z - other.z)
        if isinstance(other, Point):
            return Vec(self.x - other.x, self.y - other.y, self.z - other.z)
        raise TypeError(f"Unable to run __sub__ on a {type(self)} and a {type(other)}.")

    def to_vec(self) -> Vec:
        return Vec(*self.xyz())


class Normal(_XYZ):
    def to_vec(self) -> Vec:
        return Vec(*self.xyz())

    def squared_norm(self) -> float:
        return self.x * self.x + self.y * self.y + self.z * self.z

    def norm(self) -> float:
        return math.sqrt(self.squared_norm())

    def normalize(self) -> "Normal":
        n = self.norm()
        self.x, self.y, self.z = self.x / n, self.y / n, self.z / n
        return self


VEC_X = Vec(1.0, 0.0, 0.0)
VEC_Y = Vec(0.0, 1.0, 0.0)
VEC_Z = Vec(0.0, 0.0, 1.0)


@dataclass
class Vec2d:
    """Surface coordinates (geometry.py:233-244)."""

    u: float = 0.0
    v: float = 0.0

    def is_close(self, other: "Vec2d", epsilon: float = 1e-5) -> bool:
        return abs(self.u - other.u) < epsilon and abs(self.v - other.v) < epsilon


# ----------------------------------------------------------------------------- transformations
def _identity4() -> List[List[float]]:
    return [[1.0 if i == j else 0.0 for j in range(4)] for i in range(4)]


def _matmul4(a, b) -> List[List[float]]:
    # Accumulates k = 0..3 starting from 0.0, like transformations.py:9-16, so that composed
    # matrices carry the very same bits as the reference's.
    out = [[0.0] * 4 for _ in range(4)]
    for i in range(4):
        for j in range(4):
            acc = 0.0
            for k in range(4):
                acc += a[i][k] * b[k][j]
            out[i][j] = acc
    return out


class Transformation:
    """Affine map stored with its inverse (transformations.py:47-126).  Only ``m`` and ``invm``
    are consumed by the renderer; composition happens on the host while a scene is built."""

    def __init__(self, m=None, invm=None):
        self.m = _identity4() if m is None else m
        self.invm = _identity4() if invm is None else invm

    def __mul__(self, other):
        if isinstance(other, Transformation) or (hasattr(other, "m") and hasattr(other, "invm")):
            return Transformation(_matmul4(self.m, other.m), _matmul4(other.invm, self.invm))
        if isinstance(other, Vec):
            return Vec(*[sum(r[k] * c for k, c in enumerate(other.xyz())) for r in self.m[:3]])
        if isinstance(other, Point):
            x, y, z = [sum(r[k] * c for k, c in enumerate(other.xyz())) + r[3] for r in self.m[:3]]
            w = sum(self.m[3][k] * c for k, c in enumerate(other.xyz())) + self.m[3][3]
            return Point(x, y, z) if w == 1.0 else Point(x / w, y / w, z / w)
        if isinstance(other, Normal):
            n = other.xyz()
            return Normal(*[sum(self.invm[k][j] * n[k] for k in range(3)) for j in range(3)])
        raise TypeError(f"Invalid type {type(other)} multiplied to a Transformation object")

    def inverse(self) -> "Transformation":
        return Transformation(self.invm, self.m)

    def is_consistent(self) -> bool:
        prod = _matmul4(self.m, self.invm)
        ident = _identity4()
        return all(abs(prod[i][j] - ident[i][j]) < 1e-5 for i in range(4) for j in range(4))

    def is_close(self, other: "Transformation") -> bool:
        return all(
            abs(a[i][j] - b[i][j]) < 1e-5
            for a, b in ((self.m, other.m), (self.invm, other.invm))
            for i in range(4)
            for j in range(4)
        )


def translation(vec) -> Transformation:
    m, inv = _identity4(), _identity4()
    for i, c in enumerate((vec.x, vec.y, vec.z)):
        m[i][3], inv[i][3] = c, -c
    return Transformation(m, inv)


def scaling(vec) -> Transformation:
    m, inv = _identity4(), _identity4()
    for i, c in enumerate((vec.x, vec.y, vec.z)):
        m[i][i], inv[i][i] = c, 1 / c
    return Transformation(m, inv)


def _rotation(axis: int, angle_deg: float) -> Transformation:
    s, c = math.sin(math.radians(angle_deg)), math.cos(math.radians(angle_deg))
    i, j = [(1, 2), (2, 0), (0, 1)][axis]  # the plane that rotates, right-handed
    m, inv = _identity4(), _identity4()
    m[i][i], m[i][j], m[j][i], m[j][j] = c, -s, s, c
    inv[i][i], inv[i][j], inv[j][i], inv[j][j] = c, s, -s, c
    return Transformation(m, inv)


def rotation_x(angle_deg: float) -> Transformation:
    return _rotation(0, angle_deg)


def rotation_y(angle_deg: float) -> Transformation:
    return _rotation(1, angle_deg)


def rotation_z(angle_deg: float) -> Transformation:
    return _rotation(2, angle_deg)


# ----------------------------------------------------------------------------- rays and hits
@dataclass
class Ray:
    """ray.py:28-69"""

    origin: Point = field(default_factory=Point)
    dir: Vec = field(default_factory=Vec)
    tmin: float = 1e-5
    tmax: float = math.inf
    depth: int = 0

    def is_close(self, other: "Ray", epsilon: float = 1e-5) -> bool:
        return self.origin.is_close(other.origin, epsilon) and self.dir.is_close(other.dir, epsilon)

    def at(self, t: float) -> Point:
        return self.origin + self.dir * t

    def transform(self, transformation: Transformation) -> "Ray":
        return Ray(transformation * self.origin, transformation * self.dir, self.tmin, self.tmax, self.depth)


@dataclass
class HitRecord:
    """hitrecord.py:27-46"""

    world_point: Point
    normal: Normal
    surface_point: Vec2d
    t: float
    ray: Ray
    material: "Material" = None

    def is_close(self, other: Optional["HitRecord"], epsilon: float = 1e-5) -> bool:
        if not other:
            return False
        return (
            self.world_point.is_close(other.world_point, epsilon)
            and self.normal.is_close(other.normal, epsilon)
            and self.surface_point.is_close(other.surface_point, epsilon)
            and abs(self.t - other.t) < epsilon
            and self.ray.is_close(other.ray, epsilon)
        )


# ----------------------------------------------------------------------------- materials
class Pigment:
    """materials.py:36-47.  ``get_color`` is evaluated by the device pigment code."""

    def get_color(self, uv: Vec2d) -> Color:
        if type(self) is Pigment:
            raise NotImplementedError("Method Pigment.get_color is abstract and cannot be called")
        from . import probes

        return probes.pigment_get_color(self, uv)


class UniformPigment(Pigment):
    def __init__(self, color: Color = None):
        self.color = Color() if color is None else color


class CheckeredPigment(Pigment):
    def __init__(self, color1: Color, color2: Color, num_of_steps: int = 10):
        self.color1, self.color2, self.num_of_steps = color1, color2, num_of_steps


class ImagePigment(Pigment):
    """``image`` is anything with ``width``, ``height`` and row-major ``pixels`` (row 0 = top):
    the reference's HdrImage or :class:`pytracer_b200.hdrimage.HdrImage`."""

    def __init__(self, image):
        self.image = image


class BRDF:
    """materials.py:103-120"""

    def __init__(self, pigment: Pigment = None):
        self.pigment = UniformPigment(WHITE) if pigment is None else pigment

    def eval(self, normal: Normal, in_dir: Vec, out_dir: Vec, uv: Vec2d) -> Color:
        return BLACK

    def scatter_ray(self, pcg, incoming_dir: Vec, interaction_point: Point, normal: Normal, depth: int) -> Ray:
        if type(self) is BRDF:
            raise NotImplementedError("You cannot call BRDF.scatter_ray directly!")
        from . import probes

        return probes.brdf_scatter_ray(self, pcg, incoming_dir, interaction_point, normal, depth)


class DiffuseBRDF(BRDF):
    def eval(self, normal, in_dir, out_dir, uv) -> Color:
        return self.pigment.get_color(uv) * (1.0 / math.pi)


class SpecularBRDF(BRDF):
    def __init__(self, pigment: Pigment = None, threshold_angle_rad: float = math.pi / 1800.0):
        super().__init__(pigment)
        self.threshold_angle_rad = threshold_angle_rad


class Material:
    """materials.py:199-204"""

    def __init__(self, brdf: BRDF = None, emitted_radiance: Pigment = None):
        self.brdf = DiffuseBRDF() if brdf is None else brdf
        self.emitted_radiance = UniformPigment(BLACK) if emitted_radiance is None else emitted_radiance


# ----------------------------------------------------------------------------- shapes, lights, world
class Shape:
    """shapes.py:57-85"""

    def __init__(self, transformation: Transformation = None, material: Material = None):
        self.transformation = Transformation() if transformation is None else transformation
        self.material = Material() if material is None else material

    def _as_world(self) -> "World":
        w = World()
        w.add_shape(self)
        return w

    def ray_intersection(self, ray: Ray) -> Optional[HitRecord]:
        if type(self) is Shape:
            raise NotImplementedError("Shape.ray_intersection is an abstract method and cannot be called directly")
        return self._as_world().ray_intersection(ray, _normalize=False)

    def quick_ray_intersection(self, ray: Ray) -> bool:
        if type(self) is Shape:
            raise NotImplementedError("Shape.quick_ray_intersection is an abstract method and cannot be called directly")
        from . import probes

        return probes.quick_ray_intersection(self._as_world(), ray)


class Sphere(Shape):
    """Unit sphere at the origin, placed by ``transformation`` (shapes.py:88-151)."""


class Plane(Shape):
    """The z = 0 plane, placed by ``transformation`` (shapes.py:154-198)."""


@dataclass
class PointLight:
    """lights.py:25-39"""

    position: Point
    color: Color
    linear_radius: float = 0.0


class World:
    """world.py:26-80"""

    def __init__(self):
        self.shapes: List[Shape] = []
        self.point_lights: List[PointLight] = []

    def add_shape(self, shape: Shape) -> None:
        self.shapes.append(shape)

    def add_light(self, light: PointLight) -> None:
        self.point_lights.append(light)

    def ray_intersection(self, ray: Ray, _normalize: bool = True) -> Optional[HitRecord]:
        from . import probes

        return probes.world_ray_intersection(self, ray, normalize=_normalize)

    def is_point_visible(self, point: Point, observer_pos: Point) -> bool:
        from . import probes

        return probes.world_is_point_visible(self, point, observer_pos)


# ----------------------------------------------------------------------------- cameras
class Camera:
    """camera.py:25-39"""

    def fire_ray(self, u: float, v: float) -> Ray:
        if type(self) is Camera:
            raise NotImplementedError(f"Camera.fire_ray(u={u}, v={v}) is not implemented")
        from . import probes

        return probes.camera_fire_ray(self, u, v)


class OrthogonalCamera(Camera):
    def __init__(self, aspect_ratio: float = 1.0, transformation: Transformation = None):
        self.aspect_ratio = aspect_ratio
        self.transformation = Transformation() if transformation is None else transformation


class PerspectiveCamera(Camera):
    def __init__(self, screen_distance: float = 1.0, aspect_ratio: float = 1.0, transformation: Transformation = None):
        self.screen_distance = screen_distance
        self.aspect_ratio = aspect_ratio
        self.transformation = Transformation() if transformation is None else transformation

    def aperture_deg(self) -> float:
        return 2.0 * math.atan(self.screen_distance / self.aspect_ratio) * 180.0 / 3.14159265359


def create_onb_from_z(normal):
    """geometry.py:247-262, evaluated on the device."""
    from . import probes

    return probes.onb_from_z(normal)
