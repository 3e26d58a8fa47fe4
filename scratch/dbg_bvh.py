import sys; sys.path.insert(0, '.')
import numpy as np
from pytracer_b200 import scenes, _abi
from pytracer_b200.device import DeviceScene
from pytracer_b200.flatten import flatten_world
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG
rs = scenes.random_spheres_scene(1100, 2024, 4, 20.0, with_light=True)
sc = DeviceScene(flatten_world(rs.world))
w, h = 96, 54
for r in range(4):
    for algo in ("flat", "pointlight"):
        kw = dict(precision="f32", aa_pcg=PCG(42, 54), part_mode=_abi.RT_PART_SPP, part_rank=r, part_count=4)
        a, ha, sa = sc.render(make_params(w, h, rs.camera, algo, 2, **kw), want_hit=True)
        b, hb, sb = sc.render(make_params(w, h, rs.camera, algo, 2, accel="bvh", **kw), want_hit=True)
        d = np.argwhere((a != b).any(axis=-1))
        print(r, algo, "hit diff", int((ha != hb).sum()), "rgb diff", len(d), d[:3].tolist(), sa["rays_shadow"], sb["rays_shadow"])
        for (y, x) in d[:3]:
            print("   ", a[y, x], b[y, x], ha[y, x], hb[y, x])
