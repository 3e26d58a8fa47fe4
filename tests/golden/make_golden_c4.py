"""Oracle statistics for BASELINE config 4 on a 64x36 pixel grid (the footprints of a 60x60 box-downsample of
the 3840x2160 frame: the camera's aspect ratio comes from the scene, not from the image size).

    python tests/golden/make_golden_c4.py        # ~4 minutes on 8 host threads, writes c4_oracle_64x36.npz

16 independent renders of the config-4 scene (scenes.random_spheres_scene(1024, 2024, 4, 20.0): 1 024
ellipsoids + 2 planes, checkered / image pigments) at 64x36, 16 samples per pixel, N = 10, depth 3,
Russian roulette from 3, by the sequential fp64 C restatement of the reference (oracle/pt_oracle.c, pinned
bit for bit against the Python reference by tests/test_oracle_golden.py), rows spread over the host's
threads; run k uses the jitter stream PCG(1000 + k, 7) and the scatter stream PCG(2000 + k, 9).
The Python reference itself needs ~75 M rays x 1 026 shapes for this (weeks); the restatement 4 minutes.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import oracle  # noqa: E402
from pytracer_b200 import scenes  # noqa: E402
from pytracer_b200.flatten import flatten_world  # noqa: E402
from pytracer_b200.params import make_params  # noqa: E402
from pytracer_b200.pcg import PCG  # noqa: E402

W, H, S, RUNS, THREADS = 64, 36, 4, 16, 8  # THREADS fixes which rows share a stream (oracle.render_threaded)


def main():
    rs = scenes.random_spheres_scene(1024, 2024, 4, 20.0)
    fs = flatten_world(rs.world)
    runs, rays, samples = [], 0, 0
    for k in range(RUNS):
        p = make_params(W, H, rs.camera, algorithm="pathtracing", samples_per_side=S, num_of_rays=10, max_depth=3, rr_limit=3,
                        aa_pcg=PCG(1000 + k, 7), pt_pcg=PCG(2000 + k, 9))
        r = oracle.render_threaded(fs, p, THREADS)
        runs.append(r["rgb"].astype(np.float32))
        rays += r["rays_closest"]
        samples += r["samples"]
        print(f"run {k}: {r['rays_closest']} rays", flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c4_oracle_64x36.npz")
    np.savez_compressed(out, runs=np.stack(runs), rays_per_sample=np.float64(rays / samples), spp_per_run=np.int32(S * S), threads=np.int32(THREADS))
    print("wrote", out, "rays per sample", rays / samples)


if __name__ == "__main__":
    main()
