"""2-GPU data-parallel train step (NCCL) of both aggregators (ViS, and the ViT baseline of SURVEY §8 f-4): two ranks with half
the batch each must follow the single-GPU trajectory."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(kind):
    """(oracle module, state dict, model factory, G, B, D)."""
    if kind == "vis":
        from oracle import vis_oracle as V
        D, G, B, depth = 1024, 257, 4, 2
        sd = V.make_state_dict(1, G, input_dim=D, depth=depth)

        def make():
            from sequoia_pub_b200.tformer_lin import ViS
            return ViS(num_outputs=G, input_dim=D, depth=depth, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64)
        return V, sd, make, G, B, D
    from oracle import vit_oracle as T
    D, G, B, depth = 512, 129, 4, 2
    sd = T.make_state_dict(4, G, dim=D, depth=depth, heads=8, mlp_dim=1024)

    def make():
        from sequoia_pub_b200.vit import ViT
        return ViT(num_outputs=G, dim=D, depth=depth, heads=8, mlp_dim=1024, dim_head=64)
    return T, sd, make, G, B, D


def _worker(rank, world, port, ret, kind, comm="nccl"):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from sequoia_pub_b200.dist import split_batch
        from sequoia_pub_b200.train import FusedTrainer
        V, sd, make, G, B, D = _case(kind)
        m = make()
        m.load_state_dict(sd)
        m = m.cuda().train()
        tr = FusedTrainer(m, lr=1e-3, comm=comm)
        sl = split_batch(B, rank, world)
        try:
            for s in range(3):
                x, y = V.make_inputs(20 + s, B, G, input_dim=D)
                tr.step(x[sl].cuda(), y[sl].cuda())
        except RuntimeError as e:
            if "multicast" in str(e):
                ret[rank] = "no multicast"
                return
            raise
        x, _ = V.make_inputs(99, B, G, input_dim=D)
        with torch.no_grad():
            pred = m.eval()(x.cuda()).cpu()
        ret[rank] = pred
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,comm", [("vis", "nccl"), ("vit", "nccl"), ("vis", "multimem")])
def test_dp2_matches_single_gpu_and_oracle(kind, comm):
    """comm="multimem": the library's own NVSwitch-multicast all-reduce (csrc/comm.cu) instead of NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 2000 + (7 if kind == "vit" else 0) + (13 if comm == "multimem" else 0)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret, kind, comm), nprocs=2, join=True)
        p0, p1 = ret[0], ret[1]
    if isinstance(p0, str):
        pytest.skip("no NVSwitch multicast support on this box")
    assert torch.equal(p0, p1)                                   # replicas stay bit-identical
    V, sd, _, G, B, D = _case(kind)
    V.train_steps(sd, [V.make_inputs(20 + s, B, G, input_dim=D) for s in range(3)])
    x, _ = V.make_inputs(99, B, G, input_dim=D)
    with torch.no_grad():
        want = V.forward(sd, x)
    err = ((p0.double() - want.double()).norm() / want.double().norm()).item()
    print(f"\n[dp2 {kind}] predictions after 3 DP steps vs full-batch oracle: L2-rel {err:.3e}")
    assert err < 1e-3
