"""CPU-side checks: host logic (flatten, PCG bookkeeping, PFM, parameter marshalling) and that the
C-ABI library loads and exports every symbol include/rt_api.h declares.  No compute call is made."""
import ctypes
import io
import os
import re

import numpy as np
import pytest

from pytracer_b200 import _abi, _native, scenes
from pytracer_b200.flatten import flatten_camera, flatten_world
from pytracer_b200.hdrimage import HdrImage, read_pfm_image, InvalidPfmFileFormat
from pytracer_b200.pcg import PCG, lcg_advance
from pytracer_b200.scene import Color, Point, Transformation, Vec, rotation_x, rotation_y, rotation_z, scaling, translation
from util import demo_flat, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rt_api.h")).read()
    declared = set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.EXPORTED_SYMBOLS)
    lib = _native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rt_api_version() == _abi.RT_API_VERSION


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the same structs (gcc on the header)
    import subprocess, tempfile, textwrap

    src = textwrap.dedent("""
        #include <stdio.h>
        #include "rt_api.h"
        int main(void) {
          printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rt_pigment), sizeof(rt_material), sizeof(rt_light),
                 sizeof(rt_scene_desc), sizeof(rt_camera), sizeof(rt_render_params), sizeof(rt_stats), sizeof(rt_hit),
                 sizeof(rt_tonemap_stats));
          return 0;
        }""")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    ours = [ctypes.sizeof(t) for t in (_abi.rt_pigment, _abi.rt_material, _abi.rt_light, _abi.rt_scene_desc,
                                        _abi.rt_camera, _abi.rt_render_params, _abi.rt_stats, _abi.rt_hit,
                                        _abi.rt_tonemap_stats)]
    assert ours == sizes


def test_no_device_means_a_loud_error_not_a_fallback():
    lib = _native.load()
    if lib.rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from pytracer_b200.device import DeviceScene

    world, _ = scenes.demo_scene()
    with pytest.raises(_native.NativeError) as err:
        DeviceScene(world)
    assert err.value.code == _abi.RT_ERR_NO_DEVICE


def test_demo_scene_flattens_to_the_reference_parse():
    """scenes.demo_scene() vs the flatten of examples/demo.txt parsed by the reference's own parser."""
    z = golden("demo_scene.npz")
    world, camera = scenes.demo_scene()
    ours = flatten_world(world).to_npz_dict()
    for key, val in ours.items():
        assert np.array_equal(val, z[key]), key
    cam = flatten_camera(camera)
    assert [cam.kind, cam.screen_distance, cam.aspect_ratio] + list(cam.m) == z["camera"].tolist()


def test_pcg_host_known_answers_and_jump_ahead():
    pcg = PCG()  # tests/test_all.py:872-887
    assert pcg.state == 1753877967969059832 and pcg.inc == 109
    assert [pcg.random() for _ in range(6)] == [2707161783, 2068313097, 3122475824, 2211639955, 3215226955, 3421331566]
    a, b = PCG(45, 54), PCG(45, 54)
    for n in (0, 1, 2, 7, 1000, 123457):
        start = a.state
        for _ in range(n):
            a.random()
        b.advance(n)
        assert a.state == b.state == lcg_advance(start, a.inc, n)


def test_transformations_match_reference_known_answers():
    # tests/test_all.py:342-472 restated
    m = [[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0], [9.0, 9.0, 8.0, 7.0], [0.0, 0.0, 0.0, 1.0]]
    invm = [[-3.75, 2.75, -1, 0], [5.75, -4.75, 2.0, 1.0], [-2.25, 2.25, -1.0, -2.0], [0.0, 0.0, 0.0, 1.0]]
    t = Transformation(m, invm)
    assert t.is_consistent()
    assert (t * Vec(1.0, 2.0, 3.0)).is_close(Vec(14.0, 38.0, 51.0))
    assert (t * Point(1.0, 2.0, 3.0)).is_close(Point(18.0, 46.0, 58.0))
    from pytracer_b200.scene import Normal
    assert (t * Normal(3.0, 2.0, 4.0)).is_close(Normal(-8.75, 7.75, -3.0))
    assert (t * t.inverse()).is_close(Transformation())
    for tr in (translation(Vec(1.0, 2.0, 3.0)), scaling(Vec(2.0, 5.0, 10.0)), rotation_x(0.1), rotation_y(0.1), rotation_z(0.1)):
        assert tr.is_consistent()
    assert (rotation_x(90) * Vec(0, 1, 0)).is_close(Vec(0, 0, 1))
    assert (rotation_y(90) * Vec(0, 0, 1)).is_close(Vec(1, 0, 0))
    assert (rotation_z(90) * Vec(1, 0, 0)).is_close(Vec(0, 1, 0))
    prod = translation(Vec(1.0, 2.0, 3.0)) * translation(Vec(4.0, 6.0, 8.0))
    assert prod.is_close(translation(Vec(5.0, 8.0, 11.0)))


# tests/test_all.py:112-143: the reference's golden PFM bytes
LE_REFERENCE_BYTES = bytes([
    0x50, 0x46, 0x0a, 0x33, 0x20, 0x32, 0x0a, 0x2d, 0x31, 0x2e, 0x30, 0x0a,
    0x00, 0x00, 0xc8, 0x42, 0x00, 0x00, 0x48, 0x43, 0x00, 0x00, 0x96, 0x43,
    0x00, 0x00, 0xc8, 0x43, 0x00, 0x00, 0xfa, 0x43, 0x00, 0x00, 0x16, 0x44,
    0x00, 0x00, 0x2f, 0x44, 0x00, 0x00, 0x48, 0x44, 0x00, 0x00, 0x61, 0x44,
    0x00, 0x00, 0x20, 0x41, 0x00, 0x00, 0xa0, 0x41, 0x00, 0x00, 0xf0, 0x41,
    0x00, 0x00, 0x20, 0x42, 0x00, 0x00, 0x48, 0x42, 0x00, 0x00, 0x70, 0x42,
    0x00, 0x00, 0x8c, 0x42, 0x00, 0x00, 0xa0, 0x42, 0x00, 0x00, 0xb4, 0x42,
])


def test_pfm_write_and_read_match_the_reference_bytes():
    img = HdrImage(3, 2)
    vals = [[(1.0e1, 2.0e1, 3.0e1), (4.0e1, 5.0e1, 6.0e1), (7.0e1, 8.0e1, 9.0e1)],
            [(1.0e2, 2.0e2, 3.0e2), (4.0e2, 5.0e2, 6.0e2), (7.0e2, 8.0e2, 9.0e2)]]
    for y in range(2):
        for x in range(3):
            img.set_pixel(x, y, Color(*vals[y][x]))
    buf = io.BytesIO()
    img.write_pfm(buf)
    assert buf.getvalue() == LE_REFERENCE_BYTES
    back = read_pfm_image(io.BytesIO(LE_REFERENCE_BYTES))
    assert (back.width, back.height) == (3, 2)
    assert back.get_pixel(2, 1).is_close(Color(7.0e2, 8.0e2, 9.0e2))
    assert back.get_pixel(0, 0).is_close(Color(1.0e1, 2.0e1, 3.0e1))
    big = io.BytesIO()
    img.write_pfm(big, little_endian=False)
    assert read_pfm_image(io.BytesIO(big.getvalue())).get_pixel(1, 1).is_close(Color(4.0e2, 5.0e2, 6.0e2))
    # the reference's signature: write_pfm(stream, endianness=Endianness.LITTLE_ENDIAN), hdrimages.py:96
    from pytracer_b200 import Endianness

    for arg, want in ((Endianness.BIG_ENDIAN, big.getvalue()), (Endianness.LITTLE_ENDIAN, LE_REFERENCE_BYTES)):
        out = io.BytesIO()
        img.write_pfm(out, arg)
        assert out.getvalue() == want
        out = io.BytesIO()
        img.write_pfm(out, endianness=arg)
        assert out.getvalue() == want
    with pytest.raises(TypeError):
        img.write_pfm(io.BytesIO(), "big")
    with pytest.raises(InvalidPfmFileFormat):
        read_pfm_image(io.BytesIO(b"PF\n3 2\n-1.0\nstop"))
    assert img.pixel_offset(2, 1) == 5 and img.valid_coordinates(2, 1) and not img.valid_coordinates(3, 0)


def test_random_scene_is_reproducible_and_sized():
    a = scenes.random_spheres_scene(64, 2024, 4, 20.0)
    b = scenes.random_spheres_scene(64, 2024, 4, 20.0)
    fa, fb = flatten_world(a.world), flatten_world(b.world)
    assert fa.n_shapes == 66 and np.array_equal(fa.shape_m, fb.shape_m)
    assert len(fa.materials) == 10 and fa.texels.shape == (512 * 256, 3)
    text = a.to_scene_text()
    assert text.count("sphere(") == 64 and "e-" not in text


def test_builtin_scene_reader_matches_the_reference_parse(tmp_path, monkeypatch):
    """pytracer_b200.scene_text on the text of examples/demo.txt (restated here) == the flatten of the
    scene parsed by the reference's parser (tests/golden/demo_scene.npz)."""
    from pytracer_b200.scene_text import GrammarError, parse_scene_text

    text = '''
    float clock(150)
    material sky_material(diffuse(uniform(<0, 0, 0>)), uniform(<0.7, 0.5, 1>))
    # a comment
    material ground_material(diffuse(checkered(<0.3, 0.5, 0.1>, <0.1, 0.2, 0.5>, 4)), uniform(<0, 0, 0>))
    material sphere_material(specular(uniform(<0.5, 0.5, 0.5>)), uniform(<0, 0, 0>))
    point_light([10, 10, 10], <1, 1, 1>, 1)
    plane (sky_material, translation([0, 0, 100]) * rotation_y(clock))
    plane (ground_material, identity)
    sphere(sphere_material, translation([0, 0, 1]))
    camera(perspective, rotation_z(30) * translation([-4, 0, 1]), 1.0, 1.0)
    '''
    z = golden("demo_scene.npz")
    sc = parse_scene_text(text)
    ours = flatten_world(sc.world).to_npz_dict()
    for key, val in ours.items():
        assert np.array_equal(val, z[key]), key
    cam = flatten_camera(sc.camera)
    assert [cam.kind, cam.screen_distance, cam.aspect_ratio] + list(cam.m) == z["camera"].tolist()
    # -d NAME:VALUE overrides win over the file's own definition (scene_file.py:654-675)
    sc2 = parse_scene_text(text, {"clock": 10.0})
    assert sc2.float_variables["clock"] == 10.0
    assert not np.array_equal(flatten_world(sc2.world).shape_m, ours["shape_m"])
    # the reference's two error cases (tests/test_all.py:1309-1332)
    with pytest.raises(GrammarError):
        parse_scene_text("plane(this_material_does_not_exist, identity)")
    with pytest.raises(GrammarError):
        parse_scene_text("camera(perspective, rotation_z(30) * translation([-4, 0, 1]), 1.0, 1.0)\n"
                         "camera(orthogonal, identity, 1.0, 1.0)")
    # generated scenes round-trip through their text form
    rs = scenes.random_spheres_scene(12, 2024, 4, 20.0)
    monkeypatch.chdir(tmp_path)
    with open("texture.pfm", "wb") as f:
        rs.texture.write_pfm(f)
    back = flatten_world(parse_scene_text(rs.to_scene_text()).world).to_npz_dict()
    for key, val in flatten_world(rs.world).to_npz_dict().items():
        assert np.array_equal(val, back[key]), key


# ------------------------------------------------------------------ sphere hierarchy, host builder
def _build_bvh(m):
    m = np.ascontiguousarray(m, dtype=np.float64).reshape(-1, 12)
    n = m.shape[0]
    nodes = np.zeros((max(n, 1), 16), dtype=np.float32)
    prims = np.zeros(max(n, 1), dtype=np.int32)
    n_nodes, depth = ctypes.c_int32(0), ctypes.c_int32(0)
    lib = _native.load()
    assert lib.rt_bvh_build_host(m.ctypes.data, n, nodes.ctypes.data, nodes.shape[0], prims.ctypes.data,
                                 ctypes.byref(n_nodes), ctypes.byref(depth)) == 0, lib.rt_last_error()
    return nodes[: n_nodes.value], prims, depth.value


@pytest.mark.parametrize("n_spheres", [1, 2, 5, 64, 1024])
def test_bvh_builder_invariants_and_conservative_boxes(n_spheres):
    """The tree rt_render walks with accel = bvh, built on the host: every sphere sits in exactly one leaf
    of <= RT_BVH_LEAF spheres (rt_bvh.h), every box contains its subtree's boxes, the leaf boxes contain the ellipsoids with the
    documented padding, and a float64 walk of the tree finds every sphere a random ray really hits."""
    header = open(os.path.join(ROOT, "pytracer_b200", "csrc", "rt_bvh.h")).read()
    leaf_max = int(re.search(r"^#define RT_BVH_LEAF (\d+)", header, re.M).group(1))
    rs = scenes.random_spheres_scene(n_spheres, 11, 12, 15.0)
    fs = flatten_world(rs.world)
    sph = np.flatnonzero(fs.shape_kind == _abi.RT_SHAPE_SPHERE)
    m = fs.shape_m.reshape(-1, 12)[sph]          # spheres keep their World.shapes order inside the sorted table
    invm = fs.shape_invm.reshape(-1, 12)[sph]
    nodes, prims, depth = _build_bvh(m)
    refs = nodes[:, 12:14].copy().view(np.int32)
    boxes = nodes[:, :12].astype(np.float64).reshape(-1, 2, 2, 3)  # node, child, lo/hi, xyz
    assert depth + 2 <= 48 and len(nodes) <= max(n_spheres, 1)

    def leaf(ref):
        v = -int(ref) - 1
        return v >> 6, (v & 63) + 1

    seen = np.zeros(n_spheres, dtype=int)
    centre = m[:, [3, 7, 11]]
    half = np.sqrt(m[:, [0, 4, 8]] ** 2 + m[:, [1, 5, 9]] ** 2 + m[:, [2, 6, 10]] ** 2)

    def subtree_box(node):  # also checks containment on the way up
        out = []
        for c in range(2):
            lo, hi = boxes[node, c, 0], boxes[node, c, 1]
            ref = refs[node, c]
            if ref >= 0:
                clo, chi = subtree_box(int(ref))
            else:
                first, count = leaf(ref)
                assert count <= leaf_max
                if n_spheres <= leaf_max and c == 1:  # the never-hit second child that wraps a one-leaf scene
                    out.append((lo, hi))
                    continue
                idx = prims[first:first + count]
                seen[idx] += 1
                clo, chi = (centre[idx] - half[idx]).min(0), (centre[idx] + half[idx]).max(0)
                assert (lo < clo).all() and (hi > chi).all()  # padded: strictly outside the exact extent
            assert (lo <= clo).all() and (hi >= chi).all()
            out.append((lo, hi))
        return np.minimum(out[0][0], out[1][0]), np.maximum(out[0][1], out[1][1])

    subtree_box(0)
    assert (seen == 1).all()

    rng = np.random.default_rng(n_spheres)
    for _ in range(300):
        o = rng.uniform(-25, 25, 3); o[2] = rng.uniform(0.1, 8)
        d = rng.normal(size=3)
        # spheres the ray line really crosses (exact fp64 quadratic in the sphere's frame)
        po = invm[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].reshape(-1, 3, 3) @ o + invm[:, [3, 7, 11]]
        pd = invm[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].reshape(-1, 3, 3) @ d
        a, hb, c = (pd * pd).sum(1), (po * pd).sum(1), (po * po).sum(1) - 1
        disc = hb * hb - a * c
        t_far = (-hb + np.sqrt(np.maximum(disc, 0))) / a
        hit = set(np.flatnonzero((disc > 0) & (t_far > 0)).tolist())
        # spheres a walk of the tree reaches (slab test on the stored boxes, t >= 0)
        found, stack = set(), [0]
        inv = 1.0 / np.where(np.abs(d) < 1e-30, 1e-30, d)
        while stack:
            node = stack.pop()
            for ci in range(2):
                t0, t1 = (boxes[node, ci, 0] - o) * inv, (boxes[node, ci, 1] - o) * inv
                if np.minimum(t0, t1).max() <= np.maximum(t0, t1).min() and np.maximum(t0, t1).min() >= 0:
                    ref = refs[node, ci]
                    if ref >= 0:
                        stack.append(int(ref))
                    else:
                        first, count = leaf(ref)
                        found.update(prims[first:first + count].tolist())
        assert hit <= found


def test_tonemap_fast_path_error_stays_inside_its_guard_band():
    """k_tone_map_ldr (csrc/rt_tonemap.cu) keeps an fp32 byte only if 255*y is further than 5e-4 from an
    integer, claiming the fp32 value is within 1.4e-4 of the fp64 one.  The same arithmetic in numpy
    float32 (correctly rounded reciprocal instead of MUFU.RCP's one ulp: 255 * 2^-23 = 3e-5 of slack added)
    over 4 M values spanning 16 decades and several scales."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for scale in (1e-3, 0.37, 1.0, 563.0, 1e5):
        c = (10.0 ** rng.uniform(-8, 8, size=800_000)).astype(np.float32)
        s32 = np.float32(scale)
        s255 = np.float32(255.0) * s32
        x = c * s32
        a = c * s255
        r = np.float32(1.0) / (np.float32(1.0) + x)
        q32 = (a.astype(np.float64) * r.astype(np.float64)).astype(np.float32)  # one rounding, like the device's fma
        x64 = c.astype(np.float64) * float(np.float64(scale))
        q64 = 255.0 * (x64 / (1.0 + x64))
        worst = max(worst, float(np.abs(q32.astype(np.float64) - q64).max()))
    assert worst + 255 * 2.0 ** -23 < 1.4e-4 < 5.0e-4, worst


def test_flat_scene_knows_when_only_transformations_changed():
    """What CudaRenderer.set_world / `animate` use to decide between patching the resident scene
    (rt_scene_update_transforms) and rebuilding it."""
    from pytracer_b200.scene import Vec, translation

    a = flatten_world(scenes.demo_scene(clock=10.0)[0])
    b = flatten_world(scenes.demo_scene(clock=200.0)[0])
    assert a.differs_only_in_transforms(b) and not np.array_equal(a.shape_m, b.shape_m)
    w, _ = scenes.demo_scene(clock=10.0)
    w.shapes[2].material.brdf.pigment.color = Color(0.9, 0.1, 0.1)
    assert not a.differs_only_in_transforms(flatten_world(w))
    w, _ = scenes.demo_scene(clock=10.0)
    w.shapes.pop()
    assert not a.differs_only_in_transforms(flatten_world(w))
    w, _ = scenes.demo_scene(clock=10.0)
    w.point_lights[0].position.x += 1.0
    assert not a.differs_only_in_transforms(flatten_world(w))
    w, _ = scenes.demo_scene(clock=10.0)
    w.shapes[2].transformation = translation(Vec(0.0, 0.0, 2.0))
    assert a.differs_only_in_transforms(flatten_world(w))


def test_install_array_adopts_a_shared_frame_without_copying():
    """Multi-GPU frames land in the node's shared page-locked image; the HdrImage must take that memory
    over as it is (a 24.9 MB copy per frame was the host tail of the 8-GPU end-to-end number)."""
    from pytracer_b200.hdrimage import HdrImage, install_array

    img = HdrImage(8, 4)
    own = img.rgb_array()
    frame = np.arange(4 * 8 * 3, dtype=np.float32).reshape(4, 8, 3)
    install_array(img, frame)                      # default: the caller's buffer is kept, values copied
    assert img.rgb_array() is own and np.array_equal(own, frame)
    shared = frame * 2
    install_array(img, shared, adopt=True)         # shared frame: adopted
    assert img.rgb_array() is shared and img.get_pixel(1, 0).r == shared[0, 1, 0]
    install_array(img, shared, adopt=True)         # next frame into the same memory: nothing to do
    assert img.rgb_array() is shared
