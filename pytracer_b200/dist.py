"""One image over several GPUs: one process per GPU, ``torch.distributed`` for the single exchange.

The hot path shards without communication: every pixel sample is independent once it owns its
random stream, and the scene (a few hundred bytes to ~200 KB) is replicated.  Path tracing splits
the strata of every pixel across ranks (perfect balance: each GPU sees every pixel); the
deterministic renderers split interleaved rows (cost per row varies up to 85x on demo.txt, and 4 spp
cannot be cut 8 ways).  Each rank renders a full-size fp32 image holding only its share, already
scaled by 1/S^2, so ONE all-reduce(sum) — NCCL over NVLink on the GPUs, gloo in the CPU tests of the
host logic — leaves the finished image on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _abi


def choose_partition(algorithm: int, samples_per_side: int, world_size: int) -> int:
    """SPP split when every rank gets at least one stratum of each pixel, otherwise rows."""
    if world_size <= 1:
        return _abi.RT_PART_NONE
    spp = max(1, samples_per_side) ** 2
    if algorithm == _abi.RT_ALGO_PATHTRACING and spp >= world_size and spp % world_size == 0:
        return _abi.RT_PART_SPP
    return _abi.RT_PART_ROWS


def partition_params(params: _abi.rt_render_params, rank: int, world_size: int) -> _abi.rt_render_params:
    p = _abi.rt_render_params.from_buffer_copy(bytes(params))
    p.part_mode = choose_partition(params.algorithm, params.samples_per_side, world_size)
    p.part_rank, p.part_count = rank, world_size
    return p


def strata_of_rank(samples_per_side: int, rank: int, world_size: int):
    spp = max(1, samples_per_side) ** 2
    return list(range(rank, spp, world_size))


def rows_of_rank(height: int, rank: int, world_size: int):
    return list(range(rank, height, world_size))


@dataclass
class TorchComm:
    """Thin view of an initialised ``torch.distributed`` process group."""

    rank: int
    world_size: int
    group: object = None

    @staticmethod
    def from_env(backend: Optional[str] = None) -> "TorchComm":
        import os

        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend=backend)
        return TorchComm(dist.get_rank(), dist.get_world_size())

    def all_reduce_sum(self, tensor) -> None:
        import torch.distributed as dist

        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)

    def barrier(self) -> None:
        import torch.distributed as dist

        dist.barrier(group=self.group)


def _render_share_cuda(scene, p: _abi.rt_render_params):
    """This rank's share, rendered into a device tensor on torch's current stream."""
    import torch

    image = torch.empty((p.height, p.width, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    scene.render_device(p, image.data_ptr(), 0, stream)
    return image, scene.finish(stream)


def render_partitioned(scene, params: _abi.rt_render_params, comm, render_share=_render_share_cuda,
                       out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, dict]:
    """This rank's share on its GPU, ONE all-reduce(sum) of the fp32 image, image to the host.
    ``render_share(scene, partitioned_params) -> (tensor, stats)`` is injectable so that the exchange
    logic can be exercised on CPU (gloo) with the oracle standing in for the GPU."""
    import torch

    p = partition_params(params, comm.rank, comm.world_size)
    image, stats = render_share(scene, p)
    comm.all_reduce_sum(image)
    counts = torch.tensor([stats["rays_closest"], stats["rays_shadow"], stats["samples"]], dtype=torch.int64, device=image.device)
    comm.all_reduce_sum(counts)
    stats = dict(stats)
    stats["rays_closest"], stats["rays_shadow"], stats["samples"] = (int(v) for v in counts.tolist())
    if out is not None and out.dtype == np.float32 and out.shape == tuple(image.shape) and out.flags.c_contiguous:
        torch.from_numpy(out).copy_(image)  # straight into the caller's (page-locked) framebuffer
        return out, stats
    return image.cpu().numpy(), stats
