"""2-GPU data-parallel ViS step (NCCL): two ranks with half the batch each must follow the single-GPU trajectory."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import vis_oracle as V
        from sequoia_pub_b200.dist import split_batch
        from sequoia_pub_b200.tformer_lin import ViS
        from sequoia_pub_b200.train import FusedTrainer
        D, G, B, depth = 1024, 257, 4, 2
        sd = V.make_state_dict(1, G, input_dim=D, depth=depth)
        m = ViS(num_outputs=G, input_dim=D, depth=depth, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64)
        m.load_state_dict(sd)
        m = m.cuda().train()
        tr = FusedTrainer(m, lr=1e-3)
        sl = split_batch(B, rank, world)
        for s in range(3):
            x, y = V.make_inputs(20 + s, B, G, input_dim=D)
            tr.step(x[sl].cuda(), y[sl].cuda())
        x, _ = V.make_inputs(99, B, G, input_dim=D)
        with torch.no_grad():
            pred = m.eval()(x.cuda()).cpu()
        ret[rank] = pred
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_dp2_matches_single_gpu_and_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import vis_oracle as V
    port = 29600 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        p0, p1 = ret[0], ret[1]
    assert torch.equal(p0, p1)                                   # replicas stay bit-identical
    D, G, B, depth = 1024, 257, 4, 2
    sd = V.make_state_dict(1, G, input_dim=D, depth=depth)
    V.train_steps(sd, [V.make_inputs(20 + s, B, G, input_dim=D) for s in range(3)])
    x, _ = V.make_inputs(99, B, G, input_dim=D)
    with torch.no_grad():
        want = V.forward(sd, x)
    err = ((p0.double() - want.double()).norm() / want.double().norm()).item()
    print(f"\n[dp2] predictions after 3 DP steps vs full-batch oracle: L2-rel {err:.3e}")
    assert err < 1e-3
