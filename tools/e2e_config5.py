"""BASELINE configs[4] driver: extraction -> k-means -> data-parallel ViS training through the product API.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/e2e_config5.py --slides 256 --train-steps 10

Per rank: its share of the synthetic WSIs (4096 uint8 tiles of 256x256x3 each, rendered as a per-slide set of "tissue mode"
colours + noise so that the features carry cluster structure) goes through ResNet-50 extraction (host tiles -> features),
k-means(100) on the rank that produced the features, and the resulting [100, 2048] cluster features of all its slides form
its shard of the global batch for data-parallel ViS training (one NCCL all-reduce of the flat gradient per step, per stage).
Prints one JSON line with per-stage wall time (max over ranks), patches/s, slides/s.  Host tiles are drawn from a pool of
`--pool` distinct slides per rank (256 x 805 MB of distinct host tiles would not fit host memory)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synth_slide(seed, n=4096, modes=150):
    """uint8 [n, 256, 256, 3]: every tile is one of `modes` colours with a smooth gradient and pixel noise."""
    g = torch.Generator().manual_seed(seed)
    colours = torch.randint(30, 226, (modes, 3), generator=g).float()
    asg = torch.randint(0, modes, (n,), generator=g)
    ramp = torch.linspace(-20, 20, 256).view(1, 256, 1, 1) + torch.linspace(-10, 10, 256).view(1, 1, 256, 1)
    out = torch.empty(n, 256, 256, 3, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    for lo in range(0, n, 256):
        a = asg[lo:lo + 256]
        t = colours[a].view(-1, 1, 1, 3) + ramp + torch.randn(len(a), 256, 256, 3, generator=g) * 12
        out[lo:lo + 256] = t.clamp_(0, 255).to(torch.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slides", type=int, default=256)
    ap.add_argument("--pool", type=int, default=2)
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--genes", type=int, default=20530)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from oracle import resnet50_oracle as O            # weight generator only
    from sequoia_pub_b200.dist import shard_slides
    from sequoia_pub_b200.extract import SlideExtractor
    from sequoia_pub_b200.kmeans import KMeans
    from sequoia_pub_b200.resnet import resnet50
    from sequoia_pub_b200.tformer_lin import ViS
    from sequoia_pub_b200.train import FusedTrainer

    def sync_max(seconds):
        t = torch.tensor([seconds], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    mine = list(shard_slides(args.slides, rank, world))
    pool = [synth_slide(10_000 * rank + i) for i in range(args.pool)]
    model = resnet50().eval()
    model.load_state_dict(O.make_state_dict(0))
    model = model.to(dev)
    ex = SlideExtractor(model, 64, (256, 256), dev)
    ex(pool[0][:256])                                   # warm-up (kernel attributes, workspaces)

    barrier(); t0 = time.perf_counter()
    feats = [ex(pool[i % args.pool]) for i in range(len(mine))]
    barrier(); t_extract = sync_max(time.perf_counter() - t0)

    barrier(); t0 = time.perf_counter()
    # kmean_features.py:96-105 per slide, on the rank that produced the features (empty clusters are relocated like sklearn does)
    km = KMeans(n_clusters=100, random_state=0, device=dev)
    clusters, iters = [], []
    for f in feats:
        clusters.append(km.fit(f).cluster_features_)
        iters.append(km.n_iter_)
    barrier(); t_kmeans = sync_max(time.perf_counter() - t0)
    assert all(np.isfinite(c).all() for c in clusters), "a slide produced an empty label"

    torch.manual_seed(0)
    vis = ViS(num_outputs=args.genes, input_dim=2048, depth=6, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64,
              device=str(dev)).to(dev).train()
    x = torch.from_numpy(np.stack(clusters)).to(dev)
    y = (torch.rand(len(mine), args.genes, generator=torch.Generator().manual_seed(rank)) * 10).to(dev)
    tr = FusedTrainer(vis, lr=1e-3, weight_decay=0.0)
    first_loss = float(tr.step(x, y).item())
    for _ in range(2):
        tr.step(x, y)
    barrier(); t0 = time.perf_counter()
    for _ in range(args.train_steps):
        tr.step(x, y)
    barrier(); t_train = sync_max(time.perf_counter() - t0)

    if rank == 0:
        total = t_extract + t_kmeans + t_train
        print(json.dumps({"config": "BASELINE configs[4]", "n_gpus": world, "slides": args.slides, "patches_per_slide": 4096,
                          "extract_s": t_extract, "patches_per_s": args.slides * 4096 / t_extract,
                          "kmeans_s": t_kmeans, "kmeans_slides_per_s": args.slides / t_kmeans, "lloyd_iterations_rank0": [min(iters), max(iters)],
                          "train_steps": args.train_steps, "train_s": t_train, "train_slides_per_s": args.slides * args.train_steps / t_train,
                          "total_s": total, "first_loss_rank0": first_loss, "final_loss_rank0": float(tr.loss.item()),
                          "api": "extract.SlideExtractor -> kmeans.KMeans -> train.FusedTrainer (NCCL all-reduce of the flat gradient per backward stage)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
