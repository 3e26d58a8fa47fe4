"""The reference's own unit-level known answers (tests/test_all.py of ziotom78/pytracer), restated
against the device-backed classes of pytracer_b200: every ``ray_intersection`` / ``fire_ray`` /
``get_color`` / renderer call below is a launch of the CUDA code through the C-ABI.  Each block cites
the reference test it restates; inputs and expected values are the reference's."""
import math

import numpy as np
import pytest

from pytracer_b200 import (BLACK, WHITE, VEC_X, VEC_Y, VEC_Z, CheckeredPigment, Color, DiffuseBRDF, HdrImage,
                           ImagePigment, Material, Normal, OrthogonalCamera, PCG, PerspectiveCamera, Plane, Point,
                           Ray, Sphere, UniformPigment, Vec, Vec2d, World, create_onb_from_z, rotation_y, rotation_z,
                           scaling, translation)
from pytracer_b200.imagetracer import ImageTracer
from pytracer_b200.render import FlatRenderer, OnOffRenderer

pytestmark = pytest.mark.gpu


def hit_of(shape, origin, direction):
    return shape.ray_intersection(Ray(origin=Point(*origin), dir=Vec(*direction)))


def check_hit(hit, point, normal, uv, t):
    assert hit is not None
    assert hit.world_point.is_close(Point(*point))
    assert hit.normal.is_close(Normal(*normal))
    assert hit.surface_point.is_close(Vec2d(*uv))
    assert abs(hit.t - t) < 1e-5


# tests/test_all.py:607-705 (TestSphere.testHit / testInnerHit / testTransformation / testNormals / testNormalDirection)
def test_sphere_hits():
    sphere = Sphere()
    check_hit(hit_of(sphere, (0, 0, 2), (0, 0, -1)), (0, 0, 1), (0, 0, 1), (0.0, 0.0), 1.0)
    check_hit(hit_of(sphere, (3, 0, 0), (-1, 0, 0)), (1, 0, 0), (1, 0, 0), (0.0, 0.5), 2.0)
    assert hit_of(sphere, (0, 10, 2), (0, 0, -1)) is None
    check_hit(hit_of(sphere, (0, 0, 0), (1, 0, 0)), (1, 0, 0), (-1, 0, 0), (0.0, 0.5), 1.0)  # from inside
    moved = Sphere(transformation=translation(Vec(10.0, 0.0, 0.0)))
    check_hit(hit_of(moved, (10, 0, 2), (0, 0, -1)), (10, 0, 1), (0, 0, 1), (0.0, 0.0), 1.0)
    check_hit(hit_of(moved, (13, 0, 0), (-1, 0, 0)), (11, 0, 0), (1, 0, 0), (0.0, 0.5), 2.0)
    assert hit_of(moved, (0, 0, 2), (0, 0, -1)) is None
    assert hit_of(moved, (-10, 0, 0), (0, 0, -1)) is None
    squashed = Sphere(transformation=scaling(Vec(2.0, 1.0, 1.0)))
    n = hit_of(squashed, (1.0, 1.0, 0.0), (-1.0, -1.0, 0.0)).normal.normalize()
    assert n.is_close(Normal(1.0, 4.0, 0.0).normalize())
    mirrored = Sphere(transformation=scaling(Vec(-1.0, -1.0, -1.0)))
    n = hit_of(mirrored, (0.0, 2.0, 0.0), (0, -1, 0)).normal.normalize()
    assert n.is_close(Normal(0.0, 1.0, 0.0))


# tests/test_all.py:707-748 (TestSphere.testUVCoordinates)
@pytest.mark.parametrize("origin,direction,uv", [
    ((2.0, 0.0, 0.0), (-1, 0, 0), (0.0, 0.5)), ((0.0, 2.0, 0.0), (0, -1, 0), (0.25, 0.5)),
    ((-2.0, 0.0, 0.0), (1, 0, 0), (0.5, 0.5)), ((0.0, -2.0, 0.0), (0, 1, 0), (0.75, 0.5)),
    ((2.0, 0.0, 0.5), (-1, 0, 0), (0.0, 1 / 3)), ((2.0, 0.0, -0.5), (-1, 0, 0), (0.0, 2 / 3)),
])
def test_sphere_uv(origin, direction, uv):
    assert hit_of(Sphere(), origin, direction).surface_point.is_close(Vec2d(*uv))


# tests/test_all.py:751-819 (TestPlane)
def test_plane_hits_and_uv():
    plane = Plane()
    check_hit(hit_of(plane, (0, 0, 1), (0, 0, -1)), (0, 0, 0), (0, 0, 1), (0.0, 0.0), 1.0)
    for d in ((0, 0, 1), (1, 0, 0), (0, 1, 0)):
        assert hit_of(plane, (0, 0, 1), d) is None
    turned = Plane(transformation=rotation_y(angle_deg=90.0))
    check_hit(hit_of(turned, (1, 0, 0), (-1, 0, 0)), (0, 0, 0), (1, 0, 0), (0.0, 0.0), 1.0)
    for d in ((0, 0, 1), (1, 0, 0), (0, 1, 0)):
        assert hit_of(turned, (0, 0, 1), d) is None
    assert hit_of(plane, (0.25, 0.75, 1), (0, 0, -1)).surface_point.is_close(Vec2d(0.25, 0.75))
    assert hit_of(plane, (4.25, 7.75, 1), (0, 0, -1)).surface_point.is_close(Vec2d(0.25, 0.75))


# tests/test_all.py:822-869 (TestWorld)
def test_world_closest_hit_and_visibility():
    world = World()
    world.add_shape(Sphere(transformation=translation(VEC_X * 2)))
    world.add_shape(Sphere(transformation=translation(VEC_X * 8)))
    assert world.ray_intersection(Ray(origin=Point(0.0, 0.0, 0.0), dir=VEC_X)).world_point.is_close(Point(1.0, 0.0, 0.0))
    assert world.ray_intersection(Ray(origin=Point(10.0, 0.0, 0.0), dir=-VEC_X)).world_point.is_close(Point(9.0, 0.0, 0.0))
    origin = Point(0.0, 0.0, 0.0)
    assert not world.is_point_visible(point=Point(10.0, 0.0, 0.0), observer_pos=origin)
    assert not world.is_point_visible(point=Point(5.0, 0.0, 0.0), observer_pos=origin)
    assert world.is_point_visible(point=Point(5.0, 0.0, 0.0), observer_pos=Point(4.0, 0.0, 0.0))
    assert world.is_point_visible(point=Point(0.5, 0.0, 0.0), observer_pos=origin)
    assert world.is_point_visible(point=Point(0.0, 10.0, 0.0), observer_pos=origin)
    assert world.is_point_visible(point=Point(0.0, 0.0, 10.0), observer_pos=origin)


# tests/test_all.py:498-553 (TestCameras)
def test_cameras_corner_rays():
    for cam, parallel in ((OrthogonalCamera(aspect_ratio=2.0), True), (PerspectiveCamera(screen_distance=1.0, aspect_ratio=2.0), False)):
        rays = [cam.fire_ray(u, v) for u, v in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0))]
        if parallel:
            for r in rays[1:]:
                assert abs(rays[0].dir.cross(r.dir).squared_norm()) < 1e-10
        else:
            for r in rays[1:]:
                assert rays[0].origin.is_close(r.origin)
        for r, p in zip(rays, ((0.0, 2.0, -1.0), (0.0, -2.0, -1.0), (0.0, 2.0, 1.0), (0.0, -2.0, 1.0))):
            assert r.at(1.0).is_close(Point(*p))
    shift = translation(-VEC_Y * 2.0)
    assert OrthogonalCamera(transformation=shift * rotation_z(angle_deg=90)).fire_ray(0.5, 0.5).at(1.0).is_close(Point(0.0, -2.0, 0.0))
    assert PerspectiveCamera(transformation=shift * rotation_z(math.pi / 2.0)).fire_ray(0.5, 0.5).at(1.0).is_close(Point(0.0, -2.0, 0.0))


# tests/test_all.py:556-604 (TestImageTracer)
def test_image_tracer_orientation_coverage_and_antialiasing():
    image = HdrImage(width=4, height=2)
    tracer = ImageTracer(image=image, camera=PerspectiveCamera(aspect_ratio=2))
    assert Point(0.0, 2.0, 1.0).is_close(tracer.fire_ray(0, 0, u_pixel=0.0, v_pixel=0.0).at(1.0))
    assert Point(0.0, -2.0, -1.0).is_close(tracer.fire_ray(3, 1, u_pixel=1.0, v_pixel=1.0).at(1.0))
    assert tracer.fire_ray(0, 0, u_pixel=2.5, v_pixel=1.5).is_close(tracer.fire_ray(2, 1, u_pixel=0.5, v_pixel=0.5))
    tracer.fire_all_rays(lambda ray: Color(1.0, 2.0, 3.0))
    for row in range(image.height):
        for col in range(image.width):
            assert image.get_pixel(col, row) == Color(1.0, 2.0, 3.0)

    seen = []
    pcg = PCG()
    reference_pcg = PCG()
    small = ImageTracer(HdrImage(width=1, height=1), OrthogonalCamera(aspect_ratio=1), samples_per_side=10, pcg=pcg)

    def trace_ray(ray):
        point = ray.at(1)
        assert abs(point.x) < 1e-12 and -1.0 <= point.y <= 1.0 and -1.0 <= point.z <= 1.0
        seen.append((point.y, point.z))
        return Color(0.0, 0.0, 0.0)

    small.fire_all_rays(trace_ray)
    assert len(seen) == 100
    # stratified: sample (ir, ic) lies in its own cell of the 10x10 grid, and the generator is left
    # exactly where the reference leaves it (200 draws further)
    for k, (y, z) in enumerate(seen):
        ir, ic = divmod(k, 10)
        assert ic / 10 <= (1 - y) / 2 <= (ic + 1) / 10 and ir / 10 <= (1 - z) / 2 <= (ir + 1) / 10
    for _ in range(200):
        reference_pcg.random()
    assert pcg.state == reference_pcg.state


# tests/test_all.py:890-935 (TestPigments)
def test_pigments():
    color = Color(1.0, 2.0, 3.0)
    for uv in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0)):
        assert UniformPigment(color=color).get_color(Vec2d(*uv)).is_close(color)
    image = HdrImage(width=2, height=2)
    texels = {(0, 0): (1.0, 2.0, 3.0), (1, 0): (2.0, 3.0, 1.0), (0, 1): (2.0, 1.0, 3.0), (1, 1): (3.0, 2.0, 1.0)}
    for (x, y), c in texels.items():
        image.set_pixel(x, y, Color(*c))
    pigment = ImagePigment(image)
    for (x, y), c in texels.items():
        assert pigment.get_color(Vec2d(float(x), float(y))).is_close(Color(*c))
    c1, c2 = Color(1.0, 2.0, 3.0), Color(10.0, 20.0, 30.0)
    checker = CheckeredPigment(color1=c1, color2=c2, num_of_steps=2)
    for uv, expected in (((0.25, 0.25), c1), ((0.75, 0.25), c2), ((0.25, 0.75), c2), ((0.75, 0.75), c1)):
        assert checker.get_color(Vec2d(*uv)).is_close(expected)


# tests/test_all.py:938-988 (TestRenderers)
@pytest.mark.parametrize("cls,lit", [(OnOffRenderer, WHITE), (FlatRenderer, Color(1.0, 2.0, 3.0))])
def test_onoff_and_flat_light_only_the_centre_pixel(cls, lit):
    pigment_color = WHITE if cls is OnOffRenderer else lit
    sphere = Sphere(transformation=translation(Vec(2, 0, 0)) * scaling(Vec(0.2, 0.2, 0.2)),
                    material=Material(brdf=DiffuseBRDF(pigment=UniformPigment(pigment_color))))
    image = HdrImage(width=3, height=3)
    world = World()
    world.add_shape(sphere)
    ImageTracer(image=image, camera=OrthogonalCamera()).fire_all_rays(cls(world=world))
    for row in range(3):
        for col in range(3):
            assert image.get_pixel(col, row).is_close(lit if (col, row) == (1, 1) else BLACK)


# tests/test_all.py:991-1011 (TestOnbCreation)
def test_onb_from_random_normals_is_orthonormal():
    pcg = PCG()
    for _ in range(20):
        normal = Vec(pcg.random_float(), pcg.random_float(), pcg.random_float()).normalize()
        e1, e2, e3 = create_onb_from_z(normal)
        assert e3.is_close(normal)
        for a, b in ((e1, e2), (e2, e3), (e3, e1)):
            assert abs(a.dot(b)) < 1e-12
        for e in (e1, e2, e3):
            assert abs(e.squared_norm() - 1.0) < 1e-12


def test_abstract_methods_raise_like_the_reference():
    from pytracer_b200 import BRDF, Camera, Pigment, Shape
    from pytracer_b200.render import Renderer

    ray = Ray(origin=Point(0, 0, 0), dir=VEC_X)
    with pytest.raises(NotImplementedError):
        Shape().ray_intersection(ray)
    with pytest.raises(NotImplementedError):
        Shape().quick_ray_intersection(ray)
    with pytest.raises(NotImplementedError):
        Pigment().get_color(Vec2d(0, 0))
    with pytest.raises(NotImplementedError):
        BRDF().scatter_ray(PCG(), VEC_X, Point(0, 0, 0), Normal(0, 0, 1), 1)
    with pytest.raises(NotImplementedError):
        Camera().fire_ray(0.5, 0.5)
    with pytest.raises(NotImplementedError):
        Renderer(World())(ray)
    with pytest.raises(TypeError):
        from pytracer_b200.flatten import flatten_world

        w = World()
        w.add_shape(Shape())
        flatten_world(w)
