"""ORACLE (test infrastructure, not product code) — CPU restatement of the UNI ViT-L/16 feature extractor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

PARITY UNPINNED: the reference builds this model with the third-party `timm` package
(`timm.create_model("vit_large_patch16_224", img_size=224, patch_size=16, init_values=1e-5, num_classes=0,
dynamic_img_size=True)`, pre_processing/compute_features_hdf5.py:63-66; call `model(image[None,:])` :128), timm is neither
pinned in requirements.txt nor installed here, and the UNI weights are gated, so there is no reference output to pin
against.  This file restates timm's published VisionTransformer forward for that configuration:
patch-embed conv 16x16/16 with bias -> 196 tokens, cls token prepended, learned pos_embed added to all 197 tokens,
24 pre-LN blocks {x += ls1 * proj(softmax(q k^T / 8) v); x += ls2 * fc2(GELU(fc1(LN(x))))} with qkv bias, 16 heads x 64,
LayerNorm eps 1e-6, exact GELU, final LayerNorm, token-0 pooling, no head.  The structure is cross-checked against
`torchvision.models.vit_l_16` (same block structure without LayerScale) in tests/test_oracle_cpu.py.

Preprocessing (compute_features_hdf5.py:53-56,125-126): Resize(224) (identity for 224x224 tiles — BASELINE config 4 feeds
224x224 directly; 256x256 tiles go through PIL's antialiased bilinear resize, which is NOT restated), ToTensor (/255, CHW),
Normalize(ImageNet mean/std).
"""
import torch
import torch.nn.functional as F

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
DIM, DEPTH, HEADS, MLP, PATCH, GRID = 1024, 24, 16, 4096, 16, 14


def param_shapes(depth=DEPTH, dim=DIM, mlp=MLP):
    """timm state_dict keys and shapes (strict=True load at compute_features_hdf5.py:65-66)."""
    s = {"cls_token": (1, 1, dim), "pos_embed": (1, GRID * GRID + 1, dim), "patch_embed.proj.weight": (dim, 3, PATCH, PATCH),
         "patch_embed.proj.bias": (dim,)}
    for i in range(depth):
        b = f"blocks.{i}"
        s.update({f"{b}.norm1.weight": (dim,), f"{b}.norm1.bias": (dim,), f"{b}.attn.qkv.weight": (3 * dim, dim),
                  f"{b}.attn.qkv.bias": (3 * dim,), f"{b}.attn.proj.weight": (dim, dim), f"{b}.attn.proj.bias": (dim,),
                  f"{b}.ls1.gamma": (dim,), f"{b}.norm2.weight": (dim,), f"{b}.norm2.bias": (dim,),
                  f"{b}.mlp.fc1.weight": (mlp, dim), f"{b}.mlp.fc1.bias": (mlp,), f"{b}.mlp.fc2.weight": (dim, mlp),
                  f"{b}.mlp.fc2.bias": (dim,), f"{b}.ls2.gamma": (dim,)})
    s.update({"norm.weight": (dim,), "norm.bias": (dim,)})
    return s


def make_state_dict(seed=0, depth=DEPTH):
    """Seeded weights: trunc-normal-like 0.02 matrices, non-trivial LayerNorm affines, LayerScale gamma ~ U(0.05, 0.5)
    (UNI's learned gammas are not the 1e-5 init)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(depth).items():
        if k.endswith("gamma"):
            sd[k] = torch.rand(shp, generator=g) * 0.45 + 0.05
        elif "norm" in k and k.endswith("weight"):
            sd[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.05
        elif k in ("cls_token", "pos_embed"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        else:
            fan_in = shp[1] * (shp[2] * shp[3] if len(shp) == 4 else 1)
            sd[k] = torch.randn(shp, generator=g) * (1.0 / fan_in) ** 0.5
    return sd


def make_patches(seed, n, size=224):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)


def preprocess(patches_u8):
    """uint8 [B,224,224,3] -> float32 [B,3,224,224] (ToTensor + Normalize; Resize(224) is the identity here)."""
    x = patches_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0
    mean = torch.tensor(MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def forward(sd, x):
    """x float [B,3,224,224] -> [B,1024]; dtype follows sd / x."""
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    B = x.shape[0]
    dim = sd["cls_token"].shape[-1]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=PATCH)      # [B,1024,14,14]
    t = t.flatten(2).transpose(1, 2)                                                                # [B,196,1024]
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], dim=1) + sd["pos_embed"]
    hd = dim // HEADS
    for i in range(depth):
        b = f"blocks.{i}"
        h = F.layer_norm(t, (dim,), sd[f"{b}.norm1.weight"], sd[f"{b}.norm1.bias"], 1e-6)
        qkv = F.linear(h, sd[f"{b}.attn.qkv.weight"], sd[f"{b}.attn.qkv.bias"]).reshape(B, -1, 3, HEADS, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = torch.softmax((q * hd ** -0.5) @ k.transpose(-2, -1), dim=-1) @ v                       # [B,H,N,hd]
        a = a.transpose(1, 2).reshape(B, -1, dim)
        t = t + sd[f"{b}.ls1.gamma"] * F.linear(a, sd[f"{b}.attn.proj.weight"], sd[f"{b}.attn.proj.bias"])
        h = F.layer_norm(t, (dim,), sd[f"{b}.norm2.weight"], sd[f"{b}.norm2.bias"], 1e-6)
        h = F.linear(F.gelu(F.linear(h, sd[f"{b}.mlp.fc1.weight"], sd[f"{b}.mlp.fc1.bias"])), sd[f"{b}.mlp.fc2.weight"],
                     sd[f"{b}.mlp.fc2.bias"])
        t = t + sd[f"{b}.ls2.gamma"] * h
    t = F.layer_norm(t, (dim,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return t[:, 0]


def to_torchvision(sd):
    """The same weights under torchvision.models.vit_l_16 names (LayerScale must be 1 for the models to agree)."""
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    out = {"class_token": sd["cls_token"], "conv_proj.weight": sd["patch_embed.proj.weight"], "conv_proj.bias": sd["patch_embed.proj.bias"],
           "encoder.pos_embedding": sd["pos_embed"], "encoder.ln.weight": sd["norm.weight"], "encoder.ln.bias": sd["norm.bias"]}
    for i in range(depth):
        b, e = f"blocks.{i}", f"encoder.layers.encoder_layer_{i}"
        out.update({f"{e}.ln_1.weight": sd[f"{b}.norm1.weight"], f"{e}.ln_1.bias": sd[f"{b}.norm1.bias"],
                    f"{e}.self_attention.in_proj_weight": sd[f"{b}.attn.qkv.weight"], f"{e}.self_attention.in_proj_bias": sd[f"{b}.attn.qkv.bias"],
                    f"{e}.self_attention.out_proj.weight": sd[f"{b}.attn.proj.weight"], f"{e}.self_attention.out_proj.bias": sd[f"{b}.attn.proj.bias"],
                    f"{e}.ln_2.weight": sd[f"{b}.norm2.weight"], f"{e}.ln_2.bias": sd[f"{b}.norm2.bias"],
                    f"{e}.mlp.0.weight": sd[f"{b}.mlp.fc1.weight"], f"{e}.mlp.0.bias": sd[f"{b}.mlp.fc1.bias"],
                    f"{e}.mlp.3.weight": sd[f"{b}.mlp.fc2.weight"], f"{e}.mlp.3.bias": sd[f"{b}.mlp.fc2.bias"]})
    return out


def to_double(sd):
    return {k: v.double() for k, v in sd.items()}
