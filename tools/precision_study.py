"""What do bf16 tensor-core operands in the ResNet-50 extractor cost downstream?  

For each synthetic, structured slide (tiles = `modes` tissue colours + gradient + noise): features through
  (a) the default path (bf16 operands, fp32 accumulation), (b) the opt-in split-precision path (precision="bf16x3"),
  (c) the fp32 CPU oracle (oracle/resnet50_oracle.py, pinned to the reference class);
then k-means(100) on each feature matrix (the product KMeans; bit-exact with scikit-learn given the features) and the ViS
aggregator (seeded weights) on each set of cluster features.  Reports feature L2-rel, k-means label agreement (raw and
permutation-free: adjusted Rand index), cluster-feature and gene-prediction drift, and the throughput of both GPU modes.

    python tools/precision_study.py [--slides 3] [--tiles 512] > profiles/r02_precision_study.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synth_tiles(seed, n, modes=120):
    g = torch.Generator().manual_seed(seed)
    colours = torch.randint(30, 226, (modes, 3), generator=g).float()
    asg = torch.randint(0, modes, (n,), generator=g)
    ramp = torch.linspace(-20, 20, 256).view(1, 256, 1, 1) + torch.linspace(-10, 10, 256).view(1, 1, 256, 1)
    t = colours[asg].view(-1, 1, 1, 3) + ramp + torch.randn(n, 256, 256, 3, generator=g) * 12
    return t.clamp_(0, 255).to(torch.uint8)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slides", type=int, default=3)
    ap.add_argument("--tiles", type=int, default=512)
    ap.add_argument("--genes", type=int, default=1000)
    args = ap.parse_args()
    from sklearn.metrics import adjusted_rand_score
    from oracle import resnet50_oracle as RO
    from oracle import vis_oracle as V
    from sequoia_pub_b200.kmeans import KMeans
    from sequoia_pub_b200.resnet import resnet50
    from sequoia_pub_b200.tformer_lin import ViS
    sd = RO.make_state_dict(0)
    m = resnet50().eval(); m.load_state_dict(sd); m = m.cuda()
    vis = ViS(num_outputs=args.genes, input_dim=2048, depth=6, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64)
    vis.load_state_dict(V.make_state_dict(0, args.genes)); vis = vis.cuda().eval()
    torch.set_num_threads(os.cpu_count())
    rows = []
    for sid in range(args.slides):
        tiles = synth_tiles(100 + sid, args.tiles)
        dev = tiles.cuda()
        f_lp = torch.cat([m.extract_uint8(dev[b:b + 64]) for b in range(0, args.tiles, 64)]).cpu().numpy()
        f_hp = torch.cat([m.extract_uint8(dev[b:b + 64], precision="bf16x3") for b in range(0, args.tiles, 64)]).cpu().numpy()
        t0 = time.time()
        with torch.no_grad():
            f_ref = torch.cat([RO.forward_extract(sd, RO.preprocess(tiles[b:b + 64])) for b in range(0, args.tiles, 64)]).numpy()
        cpu_s = time.time() - t0
        out = {"slide": sid, "tiles": args.tiles, "feature_l2rel_bf16": rel(f_lp, f_ref), "feature_l2rel_bf16x3": rel(f_hp, f_ref), "oracle_cpu_s": cpu_s}
        km = {}
        for tag, f in (("bf16", f_lp), ("bf16x3", f_hp), ("fp32", f_ref)):
            k = KMeans(n_clusters=100, random_state=0).fit(f)
            km[tag] = (k.labels_.copy(), k.cluster_features_.copy(), k.seed_rows_.copy(), k.n_iter_)
        with torch.no_grad():
            preds = {tag: vis(torch.from_numpy(km[tag][1])[None].cuda()).cpu().numpy() for tag in km}
        for tag in ("bf16", "bf16x3"):
            out[f"labels_equal_{tag}"] = float((km[tag][0] == km["fp32"][0]).mean())
            out[f"labels_ari_{tag}"] = float(adjusted_rand_score(km["fp32"][0], km[tag][0]))
            out[f"seed_rows_equal_{tag}"] = float((km[tag][2] == km["fp32"][2]).mean())
            out[f"cluster_features_l2rel_{tag}"] = rel(km[tag][1], km["fp32"][1])
            out[f"gene_pred_l2rel_{tag}"] = rel(preds[tag], preds["fp32"])
            out[f"gene_pred_maxrel_{tag}"] = float(np.abs(preds[tag] - preds["fp32"]).max() / np.abs(preds["fp32"]).max())
        out["lloyd_iterations"] = {t: int(km[t][3]) for t in km}
        rows.append(out)
        print(json.dumps(out), file=sys.stderr, flush=True)
    # throughput of both modes (batch 64, device-resident tiles)
    x = torch.randint(0, 256, (64, 256, 256, 3), dtype=torch.uint8, device="cuda")
    thr = {}
    for tag, kw in (("bf16", {}), ("bf16x3", {"precision": "bf16x3"})):
        for _ in range(2):
            m.extract_uint8(x, **kw)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            m.extract_uint8(x, **kw)
        e.record(); torch.cuda.synchronize()
        thr[tag] = 64 * 5 / (s.elapsed_time(e) * 1e-3)
    print(json.dumps({"slides": rows, "patches_per_s_single_lane": thr,
                      "note": "fp32 = oracle/resnet50_oracle.py on the CPU (pinned to the reference class); k-means = product KMeans (bit-exact with scikit-learn "
                              "given the features); gene predictions = ViS (seeded weights, 1000 genes) on the cluster features of each mode"}, indent=1))


if __name__ == "__main__":
    main()
