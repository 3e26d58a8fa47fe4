import sys; sys.path.insert(0, '.')
import torch
from pytracer_b200 import scenes, _abi
from pytracer_b200.device import DeviceScene
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG
count = int(sys.argv[1])
world, camera = scenes.demo_scene()
sc = DeviceScene(world)
img = torch.empty((1080, 1920, 3), dtype=torch.float32, device='cuda')
p = make_params(1920, 1080, camera, "pathtracing", 8, num_of_rays=10, max_depth=3, rr_limit=3, aa_pcg=PCG(42, 54), pt_pcg=PCG(45, 54),
                part_mode=_abi.RT_PART_SPP if count > 1 else 0, part_rank=0, part_count=count)
sc.render_device(p, img.data_ptr()); st = sc.finish()
print(st)
