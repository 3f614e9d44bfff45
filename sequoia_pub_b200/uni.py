"""Drop-in for the UNI ViT-L/16 feature extractor the reference builds with timm
(`timm.create_model("vit_large_patch16_224", img_size=224, patch_size=16, init_values=1e-5, num_classes=0,
dynamic_img_size=True)`, pre_processing/compute_features_hdf5.py:63-66; `model(image[None,:])`, :128).

`create_model` / `VisionTransformer` keep timm's state_dict names and shapes (`cls_token`, `pos_embed`,
`patch_embed.proj.*`, `blocks.{i}.{norm1,norm2}.*`, `blocks.{i}.attn.{qkv,proj}.*`, `blocks.{i}.{ls1,ls2}.gamma`,
`blocks.{i}.mlp.{fc1,fc2}.*`, `norm.*`) so `load_state_dict(torch.load("pytorch_model.bin"), strict=True)` (:65-66) works.
The forward pass is `sq_vitl16_extract` (csrc/uni.cu): bf16 tcgen05 GEMMs with fused bias / GELU / residual epilogues,
LayerScale folded into the weights, and a fused softmax-attention kernel.  PARITY UNPINNED (no timm, no UNI weights here):
checked against oracle/uni_oracle.py, which restates timm's forward.  No CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

_NO_FWD = "sequoia_b200: only VisionTransformer.forward is implemented; sub-modules are parameter containers"


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(_NO_FWD)


class _LayerScale(_Container):
    def __init__(self, dim, init_values):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class _Attention(_Container):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(_Container):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(_Container):
    def __init__(self, dim, hidden, init_values):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.ls1 = _LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)
        self.ls2 = _LayerScale(dim, init_values)


class _PatchEmbed(_Container):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class VisionTransformer(nn.Module):
    """ViT-L/16 at 224 px, token pooling, no classification head (the configuration UNI uses)."""

    feat_type = "uni"            # dataset name prefix the extraction script writes (compute_features_hdf5.py:134-135)
    feature_dim = 1024

    def __init__(self, img_size=224, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0, init_values=1e-5,
                 num_classes=0, dynamic_img_size=True, **unused):
        super().__init__()
        if (img_size, patch_size, embed_dim, num_heads, mlp_ratio, num_classes) != (224, 16, 1024, 16, 4.0, 0) or init_values is None:
            raise NotImplementedError("sequoia_b200 implements vit_large_patch16_224 with LayerScale and num_classes=0 (UNI)")
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, 197, embed_dim) * 0.02)
        self.patch_embed = _PatchEmbed(embed_dim, patch_size)
        self.blocks = nn.Sequential(*[_Block(embed_dim, int(embed_dim * mlp_ratio), init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.depth = depth
        self._packed = None
        self._ws = None
        self._lanes = None           # [(stream, workspace)] for extract_many
        self._lanes_key = None

    def _tensors(self):
        t = [self.cls_token, self.pos_embed, self.patch_embed.proj.weight, self.patch_embed.proj.bias]
        for b in self.blocks:
            t += [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.qkv.bias, b.attn.proj.weight, b.attn.proj.bias, b.ls1.gamma,
                  b.norm2.weight, b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias, b.ls2.gamma]
        return t + [self.norm.weight, self.norm.bias]

    @_lib.with_device_of(lambda self: self.cls_token)
    def _prepack(self):
        tensors = self._tensors()
        key = (sum(t._version for t in tensors), tensors[0].data_ptr(), tensors[-1].data_ptr())
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1], self._packed[2]
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise RuntimeError("sequoia_b200 VisionTransformer runs on a B200 only: call .to('cuda') first (no CPU fallback)")
        for t in tensors:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise RuntimeError("parameters must be contiguous float32 tensors on one CUDA device")
        L = _lib.lib()
        pw = torch.zeros(L.sq_vitl16_packed_weight_elems(self.depth), dtype=torch.bfloat16, device=dev)
        pv = torch.zeros(L.sq_vitl16_packed_vec_elems(self.depth), dtype=torch.float32, device=dev)
        table = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        _lib.check(L.sq_vitl16_prepack(table, self.depth, _lib.ptr(pw), _lib.ptr(pv), _lib.stream_ptr()))
        self._packed = (key, pw, pv)
        return pw, pv

    @_lib.with_device_of(lambda self, inp, *a, **k: inp)
    def _run(self, inp, kind, batch, out=None, workspace=None):
        if self.training:
            raise RuntimeError("inference only: call .eval() (compute_features_hdf5.py:68)")
        _lib.require_device()
        pw, pv = self._prepack()
        L = _lib.lib()
        need = L.sq_vitl16_workspace_bytes(batch)
        if workspace is not None:
            if workspace.numel() < need:
                raise ValueError("workspace too small")
            ws = workspace
        else:
            if self._ws is None or self._ws.numel() < need or self._ws.device != inp.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=inp.device)
            ws = self._ws
        if out is None:
            out = torch.empty(batch, 1024, dtype=torch.float32, device=inp.device)
        _lib.check(L.sq_vitl16_extract(_lib.ptr(inp), kind, batch, self.depth, _lib.ptr(pw), _lib.ptr(pv), _lib.ptr(out), _lib.ptr(ws),
                                       ws.numel(), _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def extract_many(self, patches, out=None, batch_size=64, lanes=2):
        """All tiles of a slide resident on the device: uint8 [n,224,224,3] -> float32 [n,1024]; batches alternate between
        `lanes` CUDA streams with their own workspaces (same scheme as ResNet.extract_many)."""
        if patches.dim() != 4 or patches.shape[3] != 3 or patches.dtype != torch.uint8 or not patches.is_cuda:
            raise ValueError("extract_many expects a CUDA uint8 [n,H,W,3] tensor")
        patches = patches.contiguous()
        n, H, W = patches.shape[0], patches.shape[1], patches.shape[2]
        if out is None:
            out = torch.empty(n, 1024, dtype=torch.float32, device=patches.device)
        bs = min(batch_size, max(n, 1))
        key = (lanes, bs, H, W, str(patches.device))
        if self._lanes is None or self._lanes_key != key:
            self._lanes = [(torch.cuda.Stream(device=patches.device), self.new_lane_workspace(bs, H, W, patches.device)) for _ in range(lanes)]
            self._lanes_key = key
        self._prepack()
        main = torch.cuda.current_stream(patches.device)
        for s, _ in self._lanes:
            s.wait_stream(main)
        for i, b in enumerate(range(0, n, batch_size)):
            s, ws = self._lanes[i % lanes]
            with torch.cuda.stream(s):
                self.extract_tiles_into(patches[b:b + batch_size], out[b:b + batch_size], ws)
        for s, _ in self._lanes:
            main.wait_stream(s)
        return out

    # ------------------------------------------------------------------ hooks of the slide extractor (extract.SlideExtractor)
    def new_lane_workspace(self, batch, H, W, device):
        """Extractor workspace plus, for tiles that are not 224 px, the buffers of the `Resize(224)` step (:54)."""
        from . import preproc
        ws = {"ws": torch.empty(_lib.lib().sq_vitl16_workspace_bytes(batch), dtype=torch.uint8, device=device)}
        if (H, W) != (224, 224):
            if preproc.resize_size(H, W, 224) != (224, 224):
                raise NotImplementedError("UNI needs square tiles (Resize(224) of a non-square tile is not 224x224)")
            ws["resized"] = torch.empty(batch, 224, 224, 3, dtype=torch.uint8, device=device)
            ws["tmp"] = torch.empty(max(preproc.resize_scratch_bytes(batch, H, W, 224), 1), dtype=torch.uint8, device=device)
        return ws

    def extract_tiles_into(self, tiles, out, workspace):
        """uint8 [n,H,W,3] device tiles -> out [n,1024] on the current stream: Resize(224) (bit-identical to Pillow) when the tiles
        are not 224 px, then ToTensor / Normalize fused into the patch-embedding kernel."""
        n = tiles.shape[0]
        if tuple(tiles.shape[1:3]) != (224, 224):
            from . import preproc
            tiles = preproc.resize_tiles(tiles, 224, out=workspace["resized"][:n], tmp=workspace["tmp"])
        self._run(tiles, 0, n, out, workspace=workspace["ws"])

    @torch.no_grad()
    def forward(self, x):
        """x: float32 [B,3,224,224], normalised (the tensor `transforms_val` produces, :53-56) -> float32 [B,1024]."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224) or x.dtype != torch.float32:
            raise ValueError("expected float32 [B,3,224,224] (dynamic image sizes are not implemented)")
        return self._run(x.contiguous(), 1, x.shape[0])

    @torch.no_grad()
    def extract_uint8(self, patches, out=None):
        """patches: uint8 [B,H,W,3] raw RGB tiles -> float32 [B,1024].  224-px tiles go straight in (ToTensor + Normalize fused);
        other square sizes (the pipeline's 256-px tiles) pass through the Pillow-exact `Resize(224)` first (:54)."""
        if patches.dim() != 4 or patches.shape[3] != 3 or patches.dtype != torch.uint8:
            raise ValueError("expected uint8 [B,H,W,3]")
        patches = patches.contiguous()
        if tuple(patches.shape[1:3]) != (224, 224):
            from . import preproc
            if preproc.resize_size(patches.shape[1], patches.shape[2], 224) != (224, 224):
                raise NotImplementedError("UNI needs square tiles (Resize(224) of a non-square tile is not 224x224)")
            patches = preproc.resize_tiles(patches, 224)
        return self._run(patches, 0, patches.shape[0], out)


def create_model(name, **kwargs):
    """`timm.create_model` stand-in for the one architecture the reference requests."""
    if name != "vit_large_patch16_224":
        raise NotImplementedError(f"sequoia_b200 only provides vit_large_patch16_224 (UNI), not {name!r}")
    return VisionTransformer(**kwargs)
