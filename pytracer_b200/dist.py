"""One image over the GPUs of a node: one process per GPU, ``torch.distributed`` for the plumbing.

The hot path shards without communication: every pixel sample is independent once it owns its
random stream, and the scene (a few hundred bytes to ~200 KB) is replicated.  Two splits exist
(``rt_render_params.part_mode``):

* **rows** (default): rank r traces the image rows r, r + G, r + 2G, ... completely — interleaved,
  because the cost per row varies up to 85x on demo.txt.  Every rank runs the kernel it would run alone
  (one-pixel tasks, lane-private sums), each pixel is written by exactly one rank, and the N-GPU image
  is bit-identical to the 1-GPU image (deterministic renderers always; path tracing from 32 samples per
  pixel, where a warp's task is one pixel — below that pixels share a task, the grouping follows the
  rank's own pixel list and a pixel's fp32 sum is formed in another order: equal to rounding).  What is left to exchange is placement, not arithmetic:

  - *host image* (``render_rows_to_shared_host``, what ``fire_all_rays(..., comm=)`` does): the ranks
    share ONE page-locked host image (POSIX shared memory registered with CUDA by every rank) and each
    copies its own rows straight into it — 1/G of the frame per PCIe link, no device-side collective,
    one flag barrier in that same shared memory;
  - *device image* (``render_rows_allgather``): each rank renders its rows densely (``RT_ROWS_COMPACT``)
    into its slice of a ``[G][rows][W][3]`` tensor and one in-place NCCL all-gather of the slabs leaves
    every rank with every row;
  - *device image, no collective* (``render_rows_push``): the images of all ranks are symmetric memory
    mapped into every process and the render kernel itself stores each finished pixel into all of them
    over NVLink (``rt_render_params.peer_images``); only a barrier follows.

* **spp** (``part_mode="spp"``): rank r traces the strata s = r (mod G) of every pixel into a full-size
  image already scaled by 1/S^2 and ONE all-reduce(sum) of the fp32 image — the split BASELINE.json's
  north star names — leaves the finished image on every rank.  It costs multi-pixel tasks in the path
  tracer (5.6 % at 8 ranks) and a 25 MB reduction, so it is kept as the comparison.

gloo stands in for NCCL in the CPU tests of this host logic (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

from . import _abi


def choose_partition(algorithm: int, samples_per_side: int, world_size: int, prefer: str = "rows") -> int:
    """Rows unless the caller asks for the strata split and every rank gets a stratum of each pixel."""
    if world_size <= 1:
        return _abi.RT_PART_NONE
    spp = max(1, samples_per_side) ** 2
    if prefer == "spp" and algorithm == _abi.RT_ALGO_PATHTRACING and spp >= world_size and spp % world_size == 0:
        return _abi.RT_PART_SPP
    return _abi.RT_PART_ROWS


def partition_params(params: _abi.rt_render_params, rank: int, world_size: int, prefer: str = "rows",
                     rows_layout: int = _abi.RT_ROWS_FULL) -> _abi.rt_render_params:
    p = _abi.rt_render_params.from_buffer_copy(bytes(params))
    p.part_mode = choose_partition(params.algorithm, params.samples_per_side, world_size, prefer)
    p.part_rank, p.part_count = rank, world_size
    p.rows_layout = rows_layout if p.part_mode == _abi.RT_PART_ROWS else _abi.RT_ROWS_FULL
    return p


def strata_of_rank(samples_per_side: int, rank: int, world_size: int):
    spp = max(1, samples_per_side) ** 2
    return list(range(rank, spp, world_size))


def rows_of_rank(height: int, rank: int, world_size: int):
    return list(range(rank, height, world_size))


def rows_per_rank(height: int, world_size: int) -> int:
    """Rows of the largest share (rank 0's)."""
    return (height + world_size - 1) // world_size


class SharedHostImage:
    """One host image ``float32[H][W][3]`` shared by the ranks of a node, page-locked in every process, plus
    a small tail: per-rank counters and the flags of a barrier that lives in the same memory.

    Rank 0 creates a POSIX shared-memory segment, its name travels through ``torch.distributed`` once, every
    rank maps it and registers it with CUDA (``rt_host_register``); afterwards a frame needs no collective at
    all — each rank's ``rt_render`` copies its rows into place and bumps its flag."""

    TAIL = 4096

    def __init__(self, comm: "TorchComm", height: int, width: int, pin: bool = True):
        from multiprocessing import resource_tracker, shared_memory

        import torch.distributed as dist

        self.rank, self.world_size = comm.rank, comm.world_size
        self.height, self.width = int(height), int(width)
        nbytes = self.height * self.width * 12
        names = [None]
        if comm.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes + self.TAIL)
            names[0] = self._shm.name
        dist.broadcast_object_list(names, src=0, group=comm.group)
        if comm.rank != 0:
            self._shm = shared_memory.SharedMemory(name=names[0])
            try:  # the creator owns the segment; an attaching process must not unlink it at exit (bpo-39959)
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((self.height, self.width, 3), dtype=np.float32, buffer=self._shm.buf)
        tail = np.ndarray((self.TAIL // 8,), dtype=np.uint64, buffer=self._shm.buf, offset=nbytes)
        self.counters = tail[: 8 * 16].reshape(16, 8)   # per rank: rays_closest, rays_shadow, samples, kernel_us, ...
        self._flags = tail[8 * 16: 8 * 16 + 64]         # per rank: number of barriers passed
        if comm.rank == 0:
            tail[:] = 0
        self._passed = 0
        self.pinned = False
        if pin:
            try:
                from . import _native

                self.pinned = _native.load().rt_host_register(self.array.ctypes.data, nbytes) == 0
            except Exception:
                self.pinned = False
        comm.barrier()  # everybody has mapped the segment (and rank 0 has zeroed the tail)

    def barrier(self, timeout_s: float = 120.0) -> None:
        """All ranks of the node have reached this point (flags in the shared segment, no GPU work)."""
        self._passed += 1
        self._flags[self.rank] = self._passed
        deadline = time.monotonic() + timeout_s
        spins = 0
        while True:
            if int(self._flags[: self.world_size].min()) >= self._passed:
                return
            spins += 1
            if spins > 2000:
                time.sleep(50e-6)
                if time.monotonic() > deadline:
                    raise TimeoutError("SharedHostImage.barrier: a rank did not arrive")

    def close(self) -> None:
        try:
            if self.pinned:
                from . import _native

                _native.load().rt_host_unregister(self.array.ctypes.data)
        except Exception:
            pass
        self.pinned = False
        self.array = None
        self.counters = self._flags = None
        try:
            if self.rank == 0:  # the name goes away now; the memory when the last mapping does
                self._shm.unlink()
        except Exception:
            pass
        try:
            self._shm.close()  # (raises while an image that adopted the frame is still alive: its mapping stays)
        except Exception:
            pass


@dataclass
class TorchComm:
    """Thin view of an initialised ``torch.distributed`` process group."""

    rank: int
    world_size: int
    group: object = None
    _shared: Dict[Tuple[int, int], SharedHostImage] = field(default_factory=dict)

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

    def all_gather_slabs(self, slabs) -> None:
        """In-place all-gather of ``slabs[rank]`` into ``slabs`` ([G][...], contiguous)."""
        import torch.distributed as dist

        dist.all_gather_into_tensor(slabs, slabs[self.rank], group=self.group)

    def barrier(self) -> None:
        import torch.distributed as dist

        dist.barrier(group=self.group)

    def shared_image(self, height: int, width: int, pin: bool = True) -> SharedHostImage:
        key = (int(height), int(width))
        if key not in self._shared:
            self._shared[key] = SharedHostImage(self, height, width, pin)
        return self._shared[key]

    def close(self) -> None:
        for s in self._shared.values():
            s.close()
        self._shared.clear()


# ------------------------------------------------------------------ host image: what fire_all_rays(comm=) does
def _render_rows_to_host(scene, p: _abi.rt_render_params, out: np.ndarray) -> dict:
    """rt_render with RT_ROWS_COMPACT: the kernel on this rank's rows, then only those rows to `out`."""
    _, _, stats = scene.render(p, out=out)
    return stats


def render_rows_to_shared_host(scene, params: _abi.rt_render_params, comm: TorchComm, render_rows=_render_rows_to_host,
                               pin: bool = True) -> Tuple[np.ndarray, dict]:
    """Every rank traces its interleaved rows and copies them into the node's shared page-locked image;
    returns that image (the same memory on every rank) and the job's counters.  ``render_rows(scene,
    partitioned_params, full_size_host_array) -> stats`` is injectable so that the exchange logic can be
    exercised on CPU (gloo) with the oracle standing in for the GPU."""
    shared = comm.shared_image(params.height, params.width, pin)
    p = partition_params(params, comm.rank, comm.world_size, "rows", _abi.RT_ROWS_COMPACT)
    shared.barrier()  # nobody is still reading the previous frame out of the shared image
    stats = dict(render_rows(scene, p, shared.array))
    c = shared.counters[comm.rank]
    c[0], c[1], c[2] = stats["rays_closest"], stats["rays_shadow"], stats["samples"]
    c[3] = int(round(1e3 * stats.get("kernel_ms", 0.0)))
    shared.barrier()  # every rank's rows (and counters) are in place
    total = shared.counters[: comm.world_size]
    stats["rays_closest"], stats["rays_shadow"], stats["samples"] = (int(total[:, k].sum()) for k in range(3))
    stats["kernel_ms_max"] = float(total[:, 3].max()) / 1e3
    return shared.array, stats


# ------------------------------------------------------------------ device image
def _render_share_cuda(scene, p: _abi.rt_render_params):
    """This rank's share (RT_ROWS_FULL / strata), rendered into a device tensor on torch's current stream."""
    import torch

    image = torch.empty((p.height, p.width, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    scene.render_device(p, image.data_ptr(), 0, stream)
    return image, scene.finish(stream)


def render_partitioned(scene, params: _abi.rt_render_params, comm, render_share=_render_share_cuda,
                       out: Optional[np.ndarray] = None, prefer: str = "spp") -> Tuple[np.ndarray, dict]:
    """The all-reduce path: this rank's share as a full-size image (its strata of every pixel, or its rows
    with zeros elsewhere), ONE all-reduce(sum) of the fp32 image, image to the host on every rank.
    ``render_share(scene, partitioned_params) -> (tensor, stats)`` is injectable (CPU tests)."""
    import torch

    p = partition_params(params, comm.rank, comm.world_size, prefer)
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


class RowSlabs:
    """Device-side home of a row-split frame: ``slabs[G][rows][W][3]`` (rank r's rows, densely) and the
    full image as a strided view of it (row y = slab y % G, line y // G)."""

    def __init__(self, height: int, width: int, world_size: int):
        import torch

        self.height, self.width, self.world_size = height, width, world_size
        self.rows = rows_per_rank(height, world_size)
        self.slabs = torch.zeros((world_size, self.rows, width, 3), dtype=torch.float32, device="cuda")

    def gathered_image(self):
        """The frame as one contiguous (H, W, 3) tensor (a permuting copy of the slabs; consumers that can
        index ``slabs[y % G, y // G]`` do not need it)."""
        return self.slabs.transpose(0, 1).reshape(self.rows * self.world_size, self.width, 3)[: self.height]


def render_rows_allgather(scene, params: _abi.rt_render_params, comm: TorchComm, slabs: RowSlabs, stream: int) -> None:
    """Enqueue: this rank's rows into its slab, then the in-place all-gather.  Call scene.finish(stream)
    afterwards for the counters."""
    p = partition_params(params, comm.rank, comm.world_size, "rows", _abi.RT_ROWS_COMPACT)
    scene.render_device(p, slabs.slabs[comm.rank].data_ptr(), 0, stream)
    comm.all_gather_slabs(slabs.slabs)


class PeerImages:
    """Full-size images of all ranks as symmetric memory (``torch.distributed._symmetric_memory``): every
    process holds device pointers to every rank's image, so the render kernel can store into all of them."""

    def __init__(self, height: int, width: int, comm: TorchComm):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        group = comm.group if comm.group is not None else dist.group.WORLD
        self.image = symm.empty((height, width, 3), dtype=torch.float32, device="cuda")
        self.handle = symm.rendezvous(self.image, group.group_name if hasattr(group, "group_name") else group)
        self.pointers = [int(p) for p in self.handle.buffer_ptrs]
        assert len(self.pointers) == comm.world_size and self.pointers[comm.rank] == self.image.data_ptr()

    def barrier(self) -> None:
        self.handle.barrier()


def render_rows_push(scene, params: _abi.rt_render_params, comm: TorchComm, peers: PeerImages, stream: int) -> None:
    """Enqueue: this rank's rows, each finished pixel stored by the kernel into EVERY rank's image, then
    the symmetric-memory barrier (all stores of all ranks have landed).  No collective."""
    p = partition_params(params, comm.rank, comm.world_size, "rows", _abi.RT_ROWS_FULL)
    p.n_peer_images = comm.world_size
    for k, ptr in enumerate(peers.pointers):
        p.peer_images[k] = ptr
    scene.render_device(p, 0, 0, stream)
    peers.barrier()
