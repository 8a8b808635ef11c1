"""The N > 1 path on CPU: two gloo ranks each render their share of the frame (the CPU oracle
stands in for the per-rank renderer) and one sum-reduce assembles the image."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest

partition = importlib.import_module("path-tracing_b200.partition")
sc = conftest.pkg.scene


def test_row_block_tiles_partition_the_frame():
    for world in (1, 2, 3, 8):
        cover = np.zeros((45, 64), int)
        for rank in range(world):
            for t in partition.row_block_tiles(64, 45, rank, world, block_rows=8):
                cover[t["y0"] : t["y1"], t["x0"] : t["x1"]] += 1
        assert (cover == 1).all()
    counts = [sum(int(t["y1"] - t["y0"]) for t in partition.row_block_tiles(64, 1080, r, 8)) for r in range(8)]
    assert max(counts) - min(counts) <= 8


def test_sample_slices_tile_the_range():
    for world in (1, 2, 3, 8):
        for spp in (1, 5, 256):
            ranges = [partition.sample_slice(10, spp, r, world) for r in range(world)]
            assert ranges[0][0] == 10 and sum(c for _, c in ranges) == spp
            for (a, ca), (b, _) in zip(ranges, ranges[1:]):
                assert a + ca == b


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, conftest.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle

    scene = conftest.pkg.SceneData.load_npz(os.path.join(conftest.ROOT, "tests", "golden", "default_scene.npz"))
    o = oracle.OracleScene(scene)
    p = scene.default_params(4)
    W, H, spp = 48, 40, 4
    if mode == "tiles":
        acc, _ = o.render(p, W, H, 0, spp, tiles=partition.row_block_tiles(W, H, rank, world, block_rows=8), threads=1)
    else:
        first, count = partition.sample_slice(0, spp, rank, world)
        acc, _ = o.render(p, W, H, first, count, threads=1)
    t = torch.from_numpy(acc)
    partition.reduce_accumulation(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["tiles", "samples"])
def test_two_ranks_reduce_to_the_single_rank_image(tmp_path, default_oracle, default_scene, mode):
    out = str(tmp_path / f"{mode}.npy")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    got = np.load(out)
    ref, _ = default_oracle.render(default_scene.default_params(4), 48, 40, 0, 4)
    assert (got[..., 3] == 1).all()
    if mode == "tiles":
        assert (got == ref).all()  # disjoint pixels: bit-identical
    else:
        assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-6, atol=1e-6)  # fp32 summation order only
