"""Host-side logic of the data-parallel trainer: batch sharding, flat-buffer layout, and the gradient
all-reduce over 2 ranks with the gloo backend on CPU tensors (the N>1 path without GPUs)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from altformer_b200.trainer import FlatBuffers, GradReducer, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_batch():
    for n in (1, 7, 256, 257, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_layout_is_aligned_and_disjoint():
    shapes = [("a", (3, 22, 22)), ("b", (28,)), ("c", (768, 256)), ("d", ()), ("e", (128, 3, 1, 1))]
    lay = FlatBuffers(shapes)
    flat = torch.arange(lay.total, dtype=torch.float32)
    seen = torch.zeros(lay.total, dtype=torch.bool)
    for name, shape in shapes:
        off, n, shp = lay.index[name]
        assert off % 8 == 0 and shp == tuple(shape)
        assert not seen[off:off + n].any()
        seen[off:off + n] = True
        v = lay.view(flat, name)
        assert v.shape == torch.Size(shape) and v.data_ptr() == flat.data_ptr() + 4 * off
    assert lay.total % 8 == 0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        red = GradReducer()
        g = torch.full((1000,), float(rank + 1))
        g[rank] = 100.0
        red.reduce(g)
        mean = g * red.grad_scale
        lo, hi = shard_range(10, rank, world)
        torch.save({"mean": mean, "span": (lo, hi), "scale": red.grad_scale}, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_world2_gloo(tmp_path):
    world, port = 2, 29517 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["mean"], r1["mean"]) and r0["scale"] == 0.5
    expect = torch.full((1000,), 1.5)
    expect[0] = (100.0 + 2.0) / 2
    expect[1] = (1.0 + 100.0) / 2
    assert torch.allclose(r0["mean"], expect)
    assert r0["span"] == (0, 5) and r1["span"] == (5, 10)
