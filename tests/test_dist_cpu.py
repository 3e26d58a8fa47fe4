"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): partition choice, coverage, and the
single all-reduce(sum) exchange with the CPU oracle standing in for each rank's GPU share."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytracer_b200 import _abi
from pytracer_b200.dist import (TorchComm, choose_partition, partition_params, render_partitioned,
                                render_rows_to_shared_host, rows_of_rank, rows_per_rank, strata_of_rank)
from pytracer_b200.params import make_params
from pytracer_b200.pcg import PCG
from util import demo_flat


def test_partition_choice_and_coverage():
    PT, FLAT = _abi.RT_ALGO_PATHTRACING, _abi.RT_ALGO_FLAT
    assert choose_partition(PT, 8, 1) == _abi.RT_PART_NONE
    assert choose_partition(PT, 8, 8) == _abi.RT_PART_ROWS            # default: interleaved rows
    assert choose_partition(PT, 8, 8, "spp") == _abi.RT_PART_SPP      # 64 spp over 8 GPUs: 8 strata each
    assert choose_partition(PT, 4, 8, "spp") == _abi.RT_PART_SPP      # config 4: 16 spp, 2 strata each
    assert choose_partition(PT, 2, 8, "spp") == _abi.RT_PART_ROWS     # 4 spp cannot be cut 8 ways
    assert choose_partition(PT, 3, 2, "spp") == _abi.RT_PART_ROWS     # 9 strata do not split evenly in 2
    assert choose_partition(FLAT, 8, 4, "spp") == _abi.RT_PART_ROWS   # deterministic renderers: rows
    fs, cam = demo_flat()
    p = partition_params(make_params(48, 36, cam, "flat", 2), 3, 8, "rows", _abi.RT_ROWS_COMPACT)
    assert (p.part_mode, p.part_rank, p.part_count, p.rows_layout) == (_abi.RT_PART_ROWS, 3, 8, _abi.RT_ROWS_COMPACT)
    p = partition_params(make_params(48, 36, cam, "pathtracing", 8), 3, 8, "spp", _abi.RT_ROWS_COMPACT)
    assert (p.part_mode, p.rows_layout) == (_abi.RT_PART_SPP, _abi.RT_ROWS_FULL)
    assert rows_per_rank(1080, 8) == 135 and rows_per_rank(37, 8) == 5
    for world in (2, 4, 8):
        strata = sorted(s for r in range(world) for s in strata_of_rank(8, r, world))
        assert strata == list(range(64))
        rows = sorted(y for r in range(world) for y in rows_of_rank(1080, r, world))
        assert rows == list(range(1080))
        assert max(len(rows_of_rank(1080, r, world)) for r in range(world)) - \
            min(len(rows_of_rank(1080, r, world)) for r in range(world)) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from oracle import oracle

    comm = TorchComm(rank, world_size)
    fs, cam = demo_flat()
    params = make_params(48, 36, cam, "pointlight", 2, aa_pcg=PCG(42, 54))

    def oracle_share(flat, p):
        # rows r, r + G, ... ; the jitter stream is advanced to each row's first sample like the device does
        assert p.part_mode == _abi.RT_PART_ROWS and p.part_count == world_size and p.part_rank == rank
        rgb = np.zeros((p.height, p.width, 3))
        stats = dict(rays_closest=0, rays_shadow=0, samples=0)
        for row in rows_of_rank(p.height, rank, world_size):
            q = _abi.rt_render_params.from_buffer_copy(bytes(p))
            aa = PCG(42, 54)
            aa.advance(2 * row * p.width * 4)
            q.aa_state = aa.state
            r = oracle.render(flat, q, row, row + 1, want_hit=False, out=rgb)
            for k in stats:
                stats[k] += r[k]
        return torch.from_numpy(rgb.astype(np.float32)), stats

    image, stats = render_partitioned(fs, params, comm, render_share=oracle_share)
    if rank == 0:
        np.save(os.path.join(out_dir, "image.npy"), image)
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["rays_closest"], stats["rays_shadow"], stats["samples"]]))

    # the shared host image: every rank copies only ITS rows into one POSIX shared-memory frame (what
    # rt_render does with RT_ROWS_COMPACT), flag barriers in the same segment, counters summed through it
    def oracle_rows_to_host(flat, p, out):
        assert p.part_mode == _abi.RT_PART_ROWS and p.rows_layout == _abi.RT_ROWS_COMPACT and out.shape == (p.height, p.width, 3)
        full, st = oracle_share(flat, p)
        out[rank::world_size] = full.numpy()[rank::world_size]
        return st

    for frame in range(3):  # the segment is reused: the entry barrier keeps frames apart
        shared_img, st2 = render_rows_to_shared_host(fs, params, comm, render_rows=oracle_rows_to_host, pin=False)
        assert (st2["rays_closest"], st2["rays_shadow"], st2["samples"]) == (stats["rays_closest"], stats["rays_shadow"], stats["samples"])
        assert np.array_equal(shared_img, image), (rank, frame)   # every rank sees the whole frame
        comm.shared_image(params.height, params.width).barrier()  # everybody has compared
        shared_img[rank::world_size] = -1.0  # scribble over the own rows: the next frame must overwrite them
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_sum_to_the_single_rank_image(tmp_path):
    from oracle import oracle

    oracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    fs, cam = demo_flat()
    ref = oracle.render(fs, make_params(48, 36, cam, "pointlight", 2, aa_pcg=PCG(42, 54)), want_hit=False)
    image = np.load(tmp_path / "image.npy")
    stats = np.load(tmp_path / "stats.npy")
    assert np.array_equal(image, ref["rgb"].astype(np.float32))  # disjoint rows: x + 0 is exact
    assert stats.tolist() == [ref["rays_closest"], ref["rays_shadow"], ref["samples"]]
