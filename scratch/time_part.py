import sys; sys.path.insert(0, '.')
import torch
from pytracer_b200 import scenes, _abi
from pytracer_b200.device import DeviceScene
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG
world, camera = scenes.demo_scene()
sc = DeviceScene(world)
img = torch.empty((1080, 1920, 3), dtype=torch.float32, device='cuda')
import os
for count in [int(x) for x in os.environ.get('COUNTS','1,2,4,8,16').split(',')]:
    p = make_params(1920, 1080, camera, "pathtracing", 8, num_of_rays=10, max_depth=3, rr_limit=3, aa_pcg=PCG(42, 54), pt_pcg=PCG(45, 54),
                    part_mode=_abi.RT_PART_SPP if count > 1 else 0, part_rank=0, part_count=count)
    ms = []
    for it in range(5):
        sc.render_device(p, img.data_ptr()); st = sc.finish()
        if it >= 2: ms.append(st["kernel_ms"])
    k = sum(ms) / len(ms)
    print(f"1 of {count} ranks: kernel {k:.2f} ms, ideal {52.15 / count:.2f}, rays {st['rays_closest']}, {st['rays_closest'] / k / 1e6:.1f} Grays/s per GPU")
