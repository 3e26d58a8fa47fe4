"""Where the end-to-end time of fire_all_rays(comm=) goes at N ranks (torchrun): per-phase host-clock
times of the public path on BASELINE config 3, max / mean over ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pytracer_b200 import dist as D, _abi
from pytracer_b200.hdrimage import HdrImage
from pytracer_b200.imagetracer import CudaImageTracer
from pytracer_b200.pcg import PCG
from pytracer_b200.render import CudaRenderer

rank, local, G = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
comm = D.TorchComm.from_env("nccl") if G > 1 else None
world, camera, kw, *_ = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
himg = HdrImage(kw["width"], kw["height"])
rows = []
for it in range(8):
    t = [time.perf_counter()]
    renderer = CudaRenderer(world, algorithm=kw["algorithm"], pcg=PCG(45, 54), num_of_rays=kw.get("num_of_rays", 10),
                            max_depth=kw.get("max_depth", 10), russian_roulette_limit=kw.get("rr_limit", 3))
    tracer = CudaImageTracer(himg, camera, samples_per_side=kw["samples_per_side"], pcg=PCG(42, 54))
    torch.cuda.synchronize()
    if comm: comm.barrier()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    scene = renderer.device_scene(); t.append(time.perf_counter())          # flatten + upload
    params = tracer._params(renderer); t.append(time.perf_counter())
    if comm:
        shared = comm.shared_image(params.height, params.width, True)
        p = D.partition_params(params, comm.rank, comm.world_size, "rows", _abi.RT_ROWS_COMPACT)
        t.append(time.perf_counter())
        shared.barrier(); t.append(time.perf_counter())
        _, _, st = scene.render(p, out=shared.array); t.append(time.perf_counter())
        shared.barrier(); t.append(time.perf_counter())
        names = ["scene", "params", "partition", "barrier1", "render+d2h", "barrier2"]
    else:
        _, _, st = scene.render(params, out=tracer._adoptable_buffer()); t.append(time.perf_counter())
        names = ["scene", "params", "render+d2h"]
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])] + [st["kernel_ms"], 1e3 * (t[-1] - t[0])]
    rows.append(d)
    scene.close()
a = torch.tensor(rows[3:], dtype=torch.float64, device="cuda").mean(0)
mx, mean = a.clone(), a.clone()
if comm:
    import torch.distributed as dist
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mean); mean /= G
if rank == 0:
    for n, x, y in zip(names + ["kernel_ms", "total"], mx.tolist(), mean.tolist()):
        print(f"{n:12s} max {x:8.3f} ms   mean {y:8.3f} ms")
    print("pinned:", shared.pinned if comm else "n/a")
if comm:
    comm.close()
    import torch.distributed as dist
    dist.destroy_process_group()
