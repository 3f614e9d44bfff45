"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: slide sharding, stage-wise gradient all-reduce
over the flat buffer layout, and the equal-shard gradient identity the DP step relies on (checked with the oracle)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import vis_oracle as V
        from sequoia_pub_b200 import _lib
        from sequoia_pub_b200.dist import allreduce_stage, shard_slides, split_batch, stage_ranges
        # ---- slide sharding: every slide exactly once, balanced
        mine = list(shard_slides(11, rank, world))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert sorted(sum(gathered, [])) == list(range(11)) and max(map(len, gathered)) - min(map(len, gathered)) <= 1
        # ---- stage-wise all-reduce over the C layout == one all-reduce of the whole flat buffer
        cfg = _lib.VisConfig(256, 2, 16, 100, 37)
        ranges = stage_ranges(cfg)
        total = ranges[-1][1]
        assert ranges[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(total, generator=g)
        whole = flat.clone()
        works = [allreduce_stage(flat, ranges[s]) for s in range(cfg.depth, -1, -1)]      # head first, like the backward pass
        for w in works:
            w.wait()
        dist.all_reduce(whole)
        assert torch.equal(flat, whole)
        # ---- equal shards of a mean loss: sum of shard gradients / world == full-batch gradient
        D, G, B = 64, 9, 4
        sd = V.make_state_dict(3, G, input_dim=D, depth=1, nheads=2)
        x, y = V.make_inputs(4, B, G, input_dim=D)
        sl = split_batch(B, rank, world)
        _, _, grads = V.loss_and_grads(sd, x[sl], y[sl])
        _, _, full = V.loss_and_grads(sd, x, y)
        for k in sd:
            t = grads[k].clone()
            dist.all_reduce(t)
            assert torch.allclose(t / world, full[k], rtol=1e-4, atol=1e-7), k
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from sequoia_pub_b200.dist import shard_slides, split_batch
    assert [list(shard_slides(5, r, 3)) for r in range(3)] == [[0, 1], [2, 3], [4]]
    assert list(shard_slides(2, 3, 4)) == []
    assert split_batch(32, 1, 4) == slice(8, 16)
    with pytest.raises(ValueError):
        split_batch(10, 0, 4)
