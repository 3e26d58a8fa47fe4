"""One process driving all visible devices (rt_render_multi) on BASELINE config 3: wall time per frame
through MultiDeviceScene.render into a page-locked host image, against one device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pytracer_b200 import _native
from pytracer_b200.device import DeviceScene, MultiDeviceScene
from pytracer_b200.hdrimage import HdrImage

lib = _native.require_device()
n = lib.rt_device_count()
world, camera, kw, *_ = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
p = bench.build_params(kw, camera)
img = HdrImage(kw["width"], kw["height"]); img.pin()
out = img.rgb_array()
for g in sorted({1, 2, 4, n} & set(range(1, n + 1))):
    sc = MultiDeviceScene(world, g) if g > 1 else DeviceScene(world)
    ts = []
    for it in range(8):
        t0 = time.perf_counter()
        _, _, st = sc.render(p, out=out)
        ts.append(time.perf_counter() - t0)
    t = sum(ts[3:]) / len(ts[3:])
    rays = st["rays_closest"] + st["rays_shadow"]
    os.write(bench._REAL_STDOUT, f"{g} device(s), one process: {1e3 * t:8.3f} ms per frame end to end (slowest kernel {st['kernel_ms']:.3f} ms), {rays / t / 1e9:7.2f} Grays/s\n".encode())
    sc.close()
