"""Drop-in for the reference's `src/resnet.py` ResNet-50 feature extractor.

Same module tree and `state_dict()` keys as `resnet50()` there (torchvision names: `conv1.weight`, `bn1.*`,
`layer{1-4}.{i}.conv{1-3}.weight`, `.bn{1-3}.*`, `.downsample.{0,1}.*`, `fc.*`; src/resnet.py:96-133,370-379), so
`load_state_dict(torch.load("resnet50-19c8e357.pth"))` works unchanged.  The arithmetic of
`forward_extract` (src/resnet.py:155-170) runs in `sq_resnet50_extract` (csrc/resnet.cu): eval-mode BN is folded
into bf16 implicit-GEMM convolutions on the tcgen05 tensor cores.  There is no PyTorch fallback.

Extra entry point `extract_uint8` takes the raw uint8 HWC tiles and fuses the reference's CPU preprocessing
(`ConvertImageDtype` + `Normalize`, pre_processing/compute_features_hdf5.py:49-51,119-120) into the first kernel.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class Bottleneck(nn.Module):
    """Parameter container for one v1.5 bottleneck (stride on the 3x3; src/resnet.py:55-93)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    """ResNet-50 whose forward passes run in the sequoia_b200 CUDA library."""

    feat_type = "resnet"         # dataset name prefix the extraction script writes (compute_features_hdf5.py:134-135)
    feature_dim = 2048

    def __init__(self, block=Bottleneck, layers=(3, 4, 6, 3), num_classes=1000):
        super().__init__()
        if block is not Bottleneck or tuple(layers) != (3, 4, 6, 3):
            raise NotImplementedError("only the ResNet-50 configuration used by the SEQUOIA pipeline is implemented")
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = self._stage(64, layers[0], 1)
        self.layer2 = self._stage(128, layers[1], 2)
        self.layer3 = self._stage(256, layers[2], 2)
        self.layer4 = self._stage(512, layers[3], 2)
        self.avgpool = nn.AvgPool2d(7)
        self.fc = nn.Linear(2048, num_classes)
        # same init law as the reference (src/resnet.py:113-119): He-normal convs, BN weight 1 / bias 0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self._packed = None          # (versions, packed_w, shifts)
        self._workspace = None
        self._packed_hp = None       # (versions, hi, lo, shifts) for precision="bf16x3"
        self._workspace_hp = None
        self._lanes = None           # [(stream, workspace)] for extract_many

    def _stage(self, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes * 4))
        mods = [Bottleneck(self.inplanes, planes, stride, down)]
        self.inplanes = planes * 4
        mods += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    # ------------------------------------------------------------------ weight prepack
    def _conv_bn_pairs(self):
        pairs = [(self.conv1, self.bn1)]
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                pairs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
                if blk.downsample is not None:
                    pairs.append((blk.downsample[0], blk.downsample[1]))
        return pairs

    @_lib.with_device_of(lambda self: self.conv1.weight)
    def _prepack(self):
        tensors = []
        for conv, bn in self._conv_bn_pairs():
            tensors += [conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1], self._packed[2]
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("sequoia_b200 ResNet runs on a B200 only: call .to('cuda') first (no CPU fallback)")
        for t in tensors:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("ResNet parameters must be contiguous float32")
        L = _lib.lib()
        packed_w = torch.empty(L.sq_resnet50_packed_weight_elems(), dtype=torch.bfloat16, device=dev)
        shifts = torch.empty(L.sq_resnet50_shift_elems(), dtype=torch.float32, device=dev)
        table = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        _lib.check(L.sq_resnet50_prepack(table, _lib.ptr(packed_w), _lib.ptr(shifts), C.c_float(self.bn1.eps),
                                         _lib.stream_ptr()))
        self._packed = (key, packed_w, shifts)
        return packed_w, shifts

    @_lib.with_device_of(lambda self: self.conv1.weight)
    def _prepack_planes(self):
        """hi + lo planes of the folded weights for the split-precision mode (`precision="bf16x3"`)."""
        tensors = []
        for conv, bn in self._conv_bn_pairs():
            tensors += [conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed_hp is not None and self._packed_hp[0] == key:
            return self._packed_hp[1:]
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("sequoia_b200 ResNet runs on a B200 only: call .to('cuda') first (no CPU fallback)")
        L = _lib.lib()
        hi = torch.empty(L.sq_resnet50_packed_weight_elems(), dtype=torch.bfloat16, device=dev)
        lo = torch.empty_like(hi)
        shifts = torch.empty(L.sq_resnet50_shift_elems(), dtype=torch.float32, device=dev)
        table = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        _lib.check(L.sq_resnet50_prepack_planes(table, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(shifts), C.c_float(self.bn1.eps), _lib.stream_ptr()))
        self._packed_hp = (key, hi, lo, shifts)
        return hi, lo, shifts

    @_lib.with_device_of(lambda self, inp, *a, **k: inp)
    def _run_hp(self, inp, kind, batch, H, W, out=None):
        """Split-precision forward (hi*hi + hi*lo + lo*hi, fp32 residuals): the opt-in mode that brackets the bf16 default."""
        if self.training:
            raise RuntimeError("sequoia_b200 ResNet implements eval-mode BatchNorm only; call .eval()")
        _lib.require_device()
        hi, lo, shifts = self._prepack_planes()
        L = _lib.lib()
        need = L.sq_resnet50_hp_workspace_bytes(batch, H, W)
        if self._workspace_hp is None or self._workspace_hp.numel() < need or self._workspace_hp.device != inp.device:
            self._workspace_hp = torch.empty(need, dtype=torch.uint8, device=inp.device)
        if out is None:
            out = torch.empty(batch, 2048, dtype=torch.float32, device=inp.device)
        _lib.check(L.sq_resnet50_extract_hp(_lib.ptr(inp), kind, batch, H, W, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(shifts), _lib.ptr(out),
                                            _lib.ptr(self._workspace_hp), self._workspace_hp.numel(), _lib.stream_ptr()))
        return out

    @_lib.with_device_of(lambda self, inp, *a, **k: inp)
    def _run(self, inp, kind, batch, H, W, out=None, workspace=None):
        if self.training:
            raise RuntimeError("sequoia_b200 ResNet implements eval-mode BatchNorm only; call .eval() "
                               "(the reference does: pre_processing/compute_features_hdf5.py:60)")
        _lib.require_device()
        packed_w, shifts = self._prepack()
        L = _lib.lib()
        need = L.sq_resnet50_workspace_bytes(batch, H, W)
        if workspace is not None:
            if workspace.numel() < need:
                raise ValueError("workspace too small")
            ws = workspace
        else:
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != inp.device:
                self._workspace = torch.empty(need, dtype=torch.uint8, device=inp.device)
            ws = self._workspace
        if out is None:
            out = torch.empty(batch, 2048, dtype=torch.float32, device=inp.device)
        elif out.shape != (batch, 2048) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != inp.device:
            raise ValueError("out must be a contiguous float32 [B,2048] tensor on the input's device")
        _lib.check(L.sq_resnet50_extract(_lib.ptr(inp), kind, batch, H, W, _lib.ptr(packed_w), _lib.ptr(shifts),
                                         _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                         _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def extract_many(self, patches, out=None, batch_size=64, lanes=2):
        """All tiles of a slide resident on the device: uint8 [n,H,W,3] -> float32 [n,2048].  Batches alternate between
        `lanes` CUDA streams with their own workspaces, so the persistent convolution kernels of one batch fill the SMs
        the other batch leaves idle in partial waves, prologues and launch gaps."""
        if patches.dim() != 4 or patches.shape[3] != 3 or patches.dtype != torch.uint8 or not patches.is_cuda:
            raise ValueError("extract_many expects a CUDA uint8 [n,H,W,3] tensor")
        patches = patches.contiguous()
        n, H, W = patches.shape[0], patches.shape[1], patches.shape[2]
        if out is None:
            out = torch.empty(n, 2048, dtype=torch.float32, device=patches.device)
        need = _lib.lib().sq_resnet50_workspace_bytes(min(batch_size, max(n, 1)), H, W)
        if self._lanes is None or len(self._lanes) != lanes or self._lanes[0][1].numel() < need or self._lanes[0][1].device != patches.device:
            self._lanes = [(torch.cuda.Stream(device=patches.device), torch.empty(need, dtype=torch.uint8, device=patches.device))
                           for _ in range(lanes)]
        self._prepack()
        main = torch.cuda.current_stream(patches.device)
        for s, _ in self._lanes:
            s.wait_stream(main)
        for i, b in enumerate(range(0, n, batch_size)):
            s, ws = self._lanes[i % lanes]
            with torch.cuda.stream(s):
                self._run(patches[b:b + batch_size], 0, min(batch_size, n - b), H, W, out[b:b + batch_size], workspace=ws)
        for s, _ in self._lanes:
            main.wait_stream(s)
        return out

    # ------------------------------------------------------------------ hooks of the slide extractor (extract.SlideExtractor)
    def new_lane_workspace(self, batch, H, W, device):
        return torch.empty(_lib.lib().sq_resnet50_workspace_bytes(batch, H, W), dtype=torch.uint8, device=device)

    def extract_tiles_into(self, tiles, out, workspace):
        """uint8 [n,H,W,3] device tiles -> out [n,2048] on the current stream, using a lane's workspace."""
        self._run(tiles, 0, tiles.shape[0], tiles.shape[1], tiles.shape[2], out, workspace=workspace)

    # ------------------------------------------------------------------ reference surface
    @torch.no_grad()
    def forward_extract(self, x):
        """x: float32 [B,3,H,W], already normalised (src/resnet.py:155-170) -> float32 [B,2048]."""
        if x.dim() != 4 or x.shape[1] != 3 or x.dtype != torch.float32:
            raise ValueError("forward_extract expects float32 [B,3,H,W]")
        x = x.contiguous()
        return self._run(x, 1, x.shape[0], x.shape[2], x.shape[3])

    @torch.no_grad()
    def extract_uint8(self, patches, out=None, precision="bf16"):
        """patches: uint8 [B,H,W,3] raw RGB tiles as stored in the patch HDF5 -> float32 [B,2048].
        precision "bf16" (default): bf16 tensor-core operands, fp32 accumulation (1.4e-3 from the fp64 reference);
        "bf16x3": opt-in split-precision mode (~1e-5, about a fifth of the throughput, 256-px tiles, batches of <= 64)."""
        if patches.dim() != 4 or patches.shape[3] != 3 or patches.dtype != torch.uint8:
            raise ValueError("extract_uint8 expects uint8 [B,H,W,3]")
        patches = patches.contiguous()
        if precision == "bf16x3":
            return self._run_hp(patches, 0, patches.shape[0], patches.shape[1], patches.shape[2], out)
        if precision != "bf16":
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        return self._run(patches, 0, patches.shape[0], patches.shape[1], patches.shape[2], out)

    @torch.no_grad()
    def forward(self, x):
        """Features followed by the ImageNet classifier (src/resnet.py:138-153); fc runs as a split-precision GEMM."""
        from . import _gemm
        feat = self.forward_extract(x)
        a_hi, a_lo = _gemm.split_planes(feat)
        w_hi, w_lo = _gemm.split_planes(self.fc.weight.detach())
        out = torch.empty(feat.shape[0], self.fc.out_features, dtype=torch.float32, device=feat.device)
        _gemm.gemm(feat.shape[0], self.fc.out_features, 2048, a_hi, w_hi, a_lo, w_lo, nterms=3, out_f32=out,
                   bias=self.fc.bias.detach())
        return out


def resnet50(pretrained=False, **kwargs):
    """Same signature as the reference constructor (src/resnet.py:370-379)."""
    model = ResNet(Bottleneck, [3, 4, 6, 3], **kwargs)
    if pretrained:
        import torch.utils.model_zoo as model_zoo
        model.load_state_dict(model_zoo.load_url("https://download.pytorch.org/models/resnet50-19c8e357.pth"))
    return model
