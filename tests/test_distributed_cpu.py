"""CPU (gloo, world_size 2) tests of the data-parallel host logic: gradient buckets, averaging, batch sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vickers_hardness_unet_b200 import _lib
from vickers_hardness_unet_b200.distributed import GradBucketReducer, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ranges, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(n, generator=g)
        mine = flat.clone()
        red = GradBucketReducer(flat, ranges=ranges)
        # stage order == backward completion order; buckets are launched one by one, then waited for together
        for stage in range(len(ranges)):
            red.reduce(stage)
        # bucket-by-bucket waits (what FusedAdamW(overlap_allreduce=True) does), then finish() for whatever is left
        red.wait(0)
        red.wait(2)
        assert sorted(st for _, _, st in red._pending) == [1, 3]
        red.finish()
        assert not red._pending
        others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        want = sum(others) / world
        ok = torch.allclose(flat, want, atol=1e-6) and not torch.equal(flat, mine)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_bucket_ranges_tile_the_parameter_array_in_backward_order():
    r = _lib.grad_bucket_ranges()
    n = _lib.load().unetb200_num_params()
    assert len(r) == 4
    # stage 0 = decoder + head (end of the array), ..., stage 3 = stem + layer1 + layer2 (start of the array)
    assert r[0][1] == n and r[3][0] == 0
    assert all(r[i][0] == r[i + 1][1] for i in range(3))
    table = {name: off for name, shape, off, kind in _lib.tensor_table() if kind == 0}
    assert r[0][0] == table["decoder.blocks.0.conv1.0.weight"]
    assert r[1][0] == table["encoder.layer4.0.conv1.weight"]
    assert r[2][0] == table["encoder.layer3.0.conv1.weight"]
    assert r[1][1] - r[1][0] == 13_114_368  # encoder.layer4 (SURVEY.md section 8a census)


def test_bucketed_allreduce_averages_over_two_ranks():
    world, n = 2, 10_000
    ranges = [(7000, 10_000), (3000, 7000), (1000, 3000), (0, 1000)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, ranges, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_reducer_rejects_ranges_with_holes():
    if dist.is_initialized():
        pytest.skip("process group already initialised")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(AssertionError):
            GradBucketReducer(torch.zeros(10), ranges=[(0, 4), (5, 10)])
        GradBucketReducer(torch.zeros(10), ranges=[(4, 10), (0, 4)])
    finally:
        dist.destroy_process_group()


def test_shard_batch_partitions_every_image_once():
    for total in (1, 7, 32, 512):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
