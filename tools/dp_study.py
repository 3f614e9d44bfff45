"""Data-parallel ViS train step (BASELINE configs[2] per rank): step time per gradient-exchange mode + the all-reduce in isolation.

    torchrun --nproc-per-node N tools/dp_study.py --comm nccl|multimem [--ctas 8] [--steps 20]

Prints one JSON line (rank 0): ms/step (CUDA events, max over ranks), slides/s, per-bucket and whole-buffer all-reduce time and bus
bandwidth (2 (n-1)/n x bytes / time) measured with nothing else running."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--comm", default="nccl")
    ap.add_argument("--ctas", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import vis_oracle as V
    from sequoia_pub_b200.tformer_lin import ViS
    from sequoia_pub_b200.train import FusedTrainer
    B, G = 32, 20530
    torch.manual_seed(0)
    m = ViS(num_outputs=G, input_dim=2048, depth=6, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64, device=str(dev)).to(dev).train()
    x, y = V.make_inputs(100 + rank, B, G)
    x, y = x.to(dev), y.to(dev)
    tr = FusedTrainer(m, lr=1e-3, comm=args.comm, comm_ctas=args.ctas)

    def timed(fn, n):
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([s.elapsed_time(e) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    for _ in range(3):
        tr.step(x, y)
    ms = timed(lambda: tr.step(x, y), args.steps)
    loss = float(tr.loss.item())
    # the exchange alone, per backward stage and for the whole buffer
    rows = []
    ranges = list(tr.stage_range) + [(0, m._total)]
    for (b, e) in ranges:
        if tr._mm is not None:
            fn = lambda b=b, e=e: tr._mm.allreduce(b, e)
        else:
            fn = lambda b=b, e=e: dist.all_reduce(tr.g[b:e])
        fn()
        t = timed(fn, 10)
        nbytes = (e - b) * 4
        rows.append({"MB": nbytes / 1e6, "ms": t, "busbw_GBs": 2 * (world - 1) / world * nbytes / (t * 1e-3) / 1e9})
    if rank == 0:
        print(json.dumps({"comm": args.comm, "ctas": args.ctas if args.comm == "multimem" else None, "n_gpus": world, "ms_per_step": ms,
                          "slides_per_s": world * B / (ms * 1e-3), "loss": loss, "allreduce_per_stage": rows[:-1], "allreduce_whole_buffer": rows[-1],
                          "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_") or k.startswith("SQ_")}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
