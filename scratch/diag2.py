import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from pytracer_b200 import _abi
from pytracer_b200.device import DeviceScene
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG
from util import demo_flat
fs, cam = demo_flat()
sc = DeviceScene(fs)
base = dict(algorithm="pathtracing", samples_per_side=8, num_of_rays=10, max_depth=3, aa_pcg=PCG(42, 54), pt_pcg=PCG(45, 54))
full, _, st_full = sc.render(make_params(96, 72, cam, variant="warp", **base))
for count in (2, 4, 16, 32, 64):
    acc = np.zeros_like(full, dtype=np.float64); rays = 0
    for rank in range(count):
        part, _, st = sc.render(make_params(96, 72, cam, variant="warp", part_mode=_abi.RT_PART_SPP, part_rank=rank, part_count=count, **base))
        acc += part; rays += st["rays_closest"]
    d = np.abs(acc - full)
    rel = d / np.maximum(np.abs(full), 1e-3)
    idx = np.unravel_index(np.argmax(rel), rel.shape)
    print(count, "rays", rays, st_full["rays_closest"], "max abs", d.max(), "max rel", rel.max(), "at", idx, acc[idx], full[idx], "n>2e-5:", (rel > 2e-5).sum())
