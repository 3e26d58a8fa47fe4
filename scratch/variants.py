"""Times several builds of the library on the same workload (kernel experiments; DESIGN.md §8 cites it).

    python scratch/variants.py [--workload c3|c4|c5] [--scale K] [--accel bvh] lib1.so lib2.so ...

Each build runs in its own process (PYTRACER_B200_LIB) and prints kernel ms (mean / min of 5 after 3
warm-ups), the ray count and the image-mean luminance, so that a variant that changes the image shows."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, os
sys.path.insert(0, %(root)r)
import torch, numpy as np
import bench
bench.SCALE = %(scale)d
from pytracer_b200.device import DeviceScene
world, camera, kw, desc, fpr, n_sph = bench.workload(%(workload)r)
sc = DeviceScene(world)
p = bench.build_params(kw, camera, accel=%(accel)r, precision=%(precision)r, variant=%(variant)r)
img = torch.empty((p.height, p.width, 3), dtype=torch.float32, device='cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
ms = []
for it in range(%(iters)d):
    flush.zero_()
    sc.render_device(p, img.data_ptr()); st = sc.finish()
    if it >= %(warm)d: ms.append(st['kernel_ms'])
a = img.double()
lum = float(((a.max(-1).values + a.min(-1).values) / 2).mean())
rays = st['rays_closest'] + st['rays_shadow']
k = sum(ms) / len(ms)
os.write(bench._REAL_STDOUT, (f"{os.path.basename(os.environ.get('PYTRACER_B200_LIB', 'default')):24s} kernel {k:9.3f} ms (min {min(ms):9.3f})  rays {rays}  {rays / k / 1e6:8.3f} Grays/s  lum {lum:.6f}\n").encode())
"""


def main():
    args = sys.argv[1:]
    opt = dict(workload="c3", scale=1, accel="none", precision="auto", variant="auto", iters=8, warm=3)
    libs = []
    i = 0
    while i < len(args):
        if args[i].startswith("--"):
            key = args[i][2:]
            opt[key] = type(opt[key])(args[i + 1])
            i += 2
        else:
            libs.append(args[i])
            i += 1
    opt["root"] = ROOT
    for lib in libs or [""]:
        env = dict(os.environ)
        if lib:
            env["PYTRACER_B200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, "-c", CHILD % opt], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            sys.stdout.write(f"{lib}: FAILED\n{r.stderr[-1500:]}\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
