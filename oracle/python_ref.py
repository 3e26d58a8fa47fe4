"""TEST / BENCH INFRASTRUCTURE — times the UNMODIFIED Python reference on the host's cores.

The reference (ziotom78/pytracer) is pure Python; `pip install --no-deps --target baseline/_ref` of it
(git-ignored, travels to the GPU box with the snapshot) makes it importable there.  This
module never touches the product path: only bench.py's `cpu_baseline` / `--impl reference` legs call it.

What is timed is BASELINE config 1 exactly as BASELINE.md §3 states it — the scene of examples/demo.txt
parsed by the reference's own `parse_scene`, `ImageTracer(image, camera, samples_per_side=1)` with its
default jitter generator PCG(42, 54), `PathTracer(world, pcg=PCG(45, 54), num_of_rays=10, max_depth=3)`,
160x120 — around `tracer.fire_all_rays(renderer)` like main.py:196-198 does (process_time; wall time is
reported beside it).  Rays are counted the way BASELINE.md §3.3 counts them, by wrapping
`World.ray_intersection` and `World.is_point_visible`; the run is the reference's own sequential one, so
it must count 393 440 rays (the golden value of tests/golden/demo_c1_pathtracing_160x120.npz).
"""
from __future__ import annotations

import io
import os
import sys
import time
from pathlib import Path
from typing import Optional

_REF_DIR = Path(__file__).resolve().parent.parent / "baseline" / "_ref"

# examples/demo.txt of the reference, restated (same materials, shapes, light and camera; clock = 150)
DEMO_SCENE = """
float clock(150)
material sky_material(diffuse(uniform(<0, 0, 0>)), uniform(<0.7, 0.5, 1>))
material ground_material(diffuse(checkered(<0.3, 0.5, 0.1>, <0.1, 0.2, 0.5>, 4)), uniform(<0, 0, 0>))
material sphere_material(specular(uniform(<0.5, 0.5, 0.5>)), uniform(<0, 0, 0>))
point_light([10, 10, 10], <1, 1, 1>, 1)
plane (sky_material, translation([0, 0, 100]) * rotation_y(clock))
plane (ground_material, identity)
sphere(sphere_material, translation([0, 0, 1]))
camera(perspective, rotation_z(30) * translation([-4, 0, 1]), 1.0, 1.0)
"""
C1_RAYS = 393440


def available() -> Optional[str]:
    """None if the reference can be imported from baseline/_ref, else the reason."""
    if not (_REF_DIR / "pytracer" / "__init__.py").exists():
        return f"{_REF_DIR} does not hold the reference (pip install --no-deps --target baseline/_ref /root/reference)"
    return None


def _import_reference():
    if str(_REF_DIR) not in sys.path:
        sys.path.insert(0, str(_REF_DIR))
    import pytracer  # noqa: F401
    from pytracer import colors, hdrimages, imagetracer, pcg, render, scene_file, world

    return dict(colors=colors, hdrimages=hdrimages, imagetracer=imagetracer, pcg=pcg, render=render,
                scene_file=scene_file, world=world)


def run_config1(init_state: int = 45, init_seq: int = 54, width: int = 160, height: int = 120) -> dict:
    """One sequential run of the reference on config 1; returns rays, seconds and the image mean."""
    m = _import_reference()
    World = m["world"].World
    # Scene.world defaults to a class-level World shared by every parse in the process (scene_file.py:363):
    # give this parse a fresh one
    stream = m["scene_file"].InputStream(stream=io.StringIO(DEMO_SCENE), file_name="demo.txt")
    m["scene_file"].Scene.world = World()
    scene = m["scene_file"].parse_scene(input_file=stream, variables={})
    counts = {"closest": 0, "shadow": 0}
    ray_intersection, is_point_visible = World.ray_intersection, World.is_point_visible

    def counted_intersection(self, ray):
        counts["closest"] += 1
        return ray_intersection(self, ray)

    def counted_visible(self, point, observer_pos):
        counts["shadow"] += 1
        return is_point_visible(self, point, observer_pos)

    World.ray_intersection, World.is_point_visible = counted_intersection, counted_visible
    try:
        image = m["hdrimages"].HdrImage(width, height)
        PCG = m["pcg"].PCG
        tracer = m["imagetracer"].ImageTracer(image=image, camera=scene.camera, samples_per_side=1, pcg=PCG(42, 54))
        renderer = m["render"].PathTracer(world=scene.world, pcg=PCG(init_state=init_state, init_seq=init_seq),
                                          num_of_rays=10, max_depth=3)
        w0, c0 = time.perf_counter(), time.process_time()
        tracer.fire_all_rays(renderer)
        cpu_s, wall_s = time.process_time() - c0, time.perf_counter() - w0
    finally:
        World.ray_intersection, World.is_point_visible = ray_intersection, is_point_visible
    n = len(image.pixels)
    mean = [sum(getattr(p, ch) for p in image.pixels) / n for ch in ("r", "g", "b")]
    return dict(rays=counts["closest"] + counts["shadow"], rays_closest=counts["closest"], wall_s=wall_s, cpu_s=cpu_s,
                mean_rgb=mean, width=width, height=height)


def _worker(args):
    return run_config1(init_state=args)


def run_config1_all_cores(processes: Optional[int] = None) -> dict:
    """P independent processes, each rendering config 1 with another --init-state: aggregate rays/s
    (the reference has no parallelism of its own)."""
    import multiprocessing as mp

    processes = processes or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(processes) as pool:
        results = pool.map(_worker, [45 + k for k in range(processes)])
    wall = time.perf_counter() - t0
    rays = sum(r["rays"] for r in results)
    return dict(processes=processes, rays=rays, wall_s=wall, rays_per_s=rays / wall,
                per_process_wall_s=[r["wall_s"] for r in results])


if __name__ == "__main__":
    why = available()
    if why:
        raise SystemExit(why)
    r = run_config1()
    print(r, r["rays"] / r["wall_s"], "rays/s")
