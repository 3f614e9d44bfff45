"""Host-side data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the
B200 box, gloo in the CPU tests).

* feature extraction / k-means shard WHOLE SLIDES across ranks — the axis the reference already exposes as
  `--start/--end` row ranges (pre_processing/compute_features_hdf5.py:29-30,80-85; kmean_features.py:23-26,56-61) — with
  no data-path collective;
* ViS training splits the batch of slides evenly and sums the flat gradient buffer across ranks, one contiguous slice
  per backward stage so the exchange of the upper layers overlaps the backward pass of the lower ones.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_slides(n_slides, rank, world):
    """Indices of the slides rank `rank` processes: contiguous, balanced ranges like the reference's --start/--end."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_slides, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def stage_ranges(cfg):
    """[begin, end) element ranges of the flat parameter/gradient buffer per backward stage:
    index l < depth = transformer layer l (stage 0 also holds pos_emb1D), index depth = regression head."""
    L = _lib.lib()
    n = L.sq_vis_param_table_len(C.byref(cfg))
    if n < 0:
        _lib.check(n)
    table = (C.c_longlong * n)()
    total = C.c_longlong()
    _lib.check(L.sq_vis_param_layout(C.byref(cfg), table, n, C.byref(total)))
    depth = cfg.depth
    starts = [table[1 + 18 * l] for l in range(depth)] + [table[n - 4], total.value]
    return [(0 if l == 0 else starts[l], starts[l + 1]) for l in range(depth)] + [(starts[depth], total.value)]


def allreduce_stage(flat, rng, group=None):
    """Starts the sum all-reduce of one stage's slice; returns the async work handle."""
    b, e = rng
    return dist.all_reduce(flat[b:e], op=dist.ReduceOp.SUM, group=group, async_op=True)


def split_batch(batch, rank, world):
    """Equal shards of a global batch (MSELoss is a mean, so equal shards + grad/world == large-batch gradient)."""
    if batch % world != 0:
        raise ValueError(f"global batch {batch} is not divisible by world size {world}")
    per = batch // world
    return slice(rank * per, (rank + 1) * per)


class MultimemAllReduce:
    """Flat fp32 gradient buffer in symmetric memory with an NVSwitch multicast mapping + the in-place all-reduce over it
    (`sq_multimem_allreduce_f32`, csrc/comm.cu: one multimem.ld_reduce and one multimem.st per 16 bytes, a handful of CTAs).
    torch.distributed._symmetric_memory only allocates and exchanges the handles (plumbing); raises RuntimeError when the
    platform has no multicast support (then the trainer keeps NCCL)."""

    def __init__(self, numel, device, group=None, ctas=8):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ctas = int(ctas)
        self.buf = symm.empty(numel, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        mc = int(self.hdl.multicast_ptr or 0)
        if mc == 0:
            raise RuntimeError("sequoia_b200: no NVSwitch multicast support for symmetric memory on this system")
        if int(getattr(self.hdl, "offset", 0) or 0) != 0:
            raise RuntimeError("sequoia_b200: symmetric gradient buffer is not at the start of its multicast allocation")
        self.mc_ptr = mc
        nflag = _lib.lib().sq_multimem_flag_bytes(self.ctas) // 4
        self.flags = symm.empty(nflag, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.fhdl = symm.rendezvous(self.flags, group)
        if int(getattr(self.fhdl, "offset", 0) or 0) != 0:
            raise RuntimeError("sequoia_b200: symmetric flag buffer is not at the start of its allocation")
        ptrs = [int(p) for p in self.fhdl.buffer_ptrs]
        self._peer_flags = (C.c_void_p * self.world)(*ptrs)
        self.epoch = 1
        torch.cuda.synchronize(device)
        dist.barrier(group)                      # every rank's flags are zero before the first kernel signals

    def allreduce(self, begin, end):
        """Enqueues the all-reduce of buf[begin:end] on the current stream (every rank must issue the same sequence)."""
        _lib.check(_lib.lib().sq_multimem_allreduce_f32(C.c_void_p(self.mc_ptr), begin, end - begin, self._peer_flags, self.rank, self.world,
                                                        self.epoch, self.ctas, _lib.stream_ptr(self.buf)))
        self.epoch += 2
