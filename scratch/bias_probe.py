"""Which kernel / precision / accel agrees with the oracle's image mean on a many-sphere scene?
(scratch: run on the GPU box; the oracle references are made on CPU by the scripts quoted in DESIGN.md)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pytracer_b200 import scenes
from pytracer_b200.device import DeviceScene
from pytracer_b200.flatten import flatten_world
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG

lum = lambda a: ((a.max(-1) + a.min(-1)) / 2)

def probe(name, rs, ref_file, w, h, S, args, seeds=8):
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), ref_file))
    runs = ref["runs"]
    l = np.array([lum(x).mean() for x in runs])
    print(f"== {name}: oracle lum {l.mean():.5f} +- {l.std(ddof=1) / np.sqrt(len(l)):.5f}, rays/sample {float(ref['rays_per_sample']):.3f}")
    sc = DeviceScene(flatten_world(rs.world))
    for variant, prec, accel in (("warp", "f32", "none"), ("warp", "f32", "bvh"), ("mega", "f32", "none"), ("mega", "f32", "bvh"), ("mega", "f64", "none")):
        vals, rps = [], []
        for k in range(seeds):
            rgb, _, st = sc.render(make_params(w, h, rs.camera, samples_per_side=S, aa_pcg=PCG(11 + k, 3), pt_pcg=PCG(77 + k, 5),
                                               variant=variant, precision=prec, accel=accel, **args))
            vals.append(lum(rgb.astype(np.float64)).mean())
            rps.append(st["rays_closest"] / st["samples"])
        vals = np.array(vals)
        print(f"{variant:5s} {prec} {accel:4s}: lum {vals.mean():.5f} +- {vals.std(ddof=1) / np.sqrt(seeds):.5f} "
              f"({(vals.mean() / l.mean() - 1) * 100:+.3f} %), rays/sample {np.mean(rps):.3f}, {st['kernel_ms']:.1f} ms")

rs = scenes.random_spheres_scene(1100, 2024, 4, 20.0, with_light=True)
probe("1100 spheres 32x18 N=4 depth 2", rs, "ref_1100_32x18.npz", 32, 18, 32, dict(algorithm="pathtracing", num_of_rays=4, max_depth=2))
if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_c4_64x36.npz")):
    rs = scenes.random_spheres_scene(1024, 2024, 4, 20.0)
    probe("config 4 scene 64x36 N=10 depth 3", rs, "ref_c4_64x36.npz", 64, 36, 16, dict(algorithm="pathtracing", num_of_rays=10, max_depth=3, rr_limit=3), seeds=4)
