"""pytracer_b200 — B200-native (sm_100a) implementation of pytracer's per-pixel rendering hot path.

``ImageTracer.fire_all_rays`` driving the OnOff / Flat / PointLight / PathTracer renderers of
ziotom78/pytracer, rebuilt as hand-written CUDA kernels behind the reference's own
``Renderer`` / ``ImageTracer`` interface.  See DESIGN.md and INTEGRATION.md.
"""
from .scene import (  # noqa: F401
    BLACK, WHITE, VEC_X, VEC_Y, VEC_Z, BRDF, Camera, CheckeredPigment, Color, DiffuseBRDF, HitRecord,
    ImagePigment, Material, Normal, OrthogonalCamera, PerspectiveCamera, Pigment, Plane, Point,
    PointLight, Ray, Shape, SpecularBRDF, Sphere, Transformation, UniformPigment, Vec, Vec2d, World,
    create_onb_from_z, rotation_x, rotation_y, rotation_z, scaling, translation,
)
from .pcg import PCG  # noqa: F401
from .hdrimage import Endianness, HdrImage, read_pfm_image  # noqa: F401

__version__ = "0.1.0"
