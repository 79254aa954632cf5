"""Data-parallel gradient exchange: bucketed all-reduce overlapped with backward.

The reference is single-device; BASELINE.json's north_star adds plain data parallelism: the batch
is sharded across ranks, BatchNorm statistics stay local per rank (ordinary nn.BatchNorm2d semantics), weights and
optimizer state are replicated and the only exchange step is the average of the gradients.

Parameters are grouped into flat fp32 buckets in reverse registration order (the order gradients become ready:
head first, stem last).  A post-accumulate-grad hook packs each gradient into its bucket and, when the bucket is
complete, launches an asynchronous all-reduce (NCCL over NVLink on the GPUs, gloo in the CPU tests) that overlaps
the rest of backward.  `finish()` waits and writes the averaged gradients back.

Zero-copy path for the convolution weights (97 % of the bytes): each 4-D weight's bucket slice is registered as its
gradient slot (kernels.register_grad_slot); the tcgen05 weight-gradient kernel writes its result straight into the
bucket and autograd adopts a view of it as `weight.grad`, so nothing is packed before the all-reduce and nothing is
copied back after it — the all-reduce averages `weight.grad` in place.
"""
from __future__ import annotations

import contextlib
from typing import List

import torch
import torch.distributed as dist


def _wgrad_side_stream():
    from .kernels import wgrad_side_stream
    return wgrad_side_stream()


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter], device):
        self.params = params
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + 63) // 64 * 64      # 256 B aligned slices: kernels write gradients straight into them
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.pending = len(params)
        self.work = None

    def view(self, i):
        p = self.params[i]
        return self.flat[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)


class GradReducer:
    """Usage:  reducer = GradReducer(model); ...; loss.backward(); reducer.finish(); optimizer.step()"""

    def __init__(self, model: torch.nn.Module, bucket_mb: float = 32.0, group=None, broadcast: bool = True,
                 zero_copy: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("GradReducer needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.avg_native = dist.get_backend(group) == "nccl"
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        if broadcast and self.world > 1:
            # replicas start identical whatever each rank's seed was (torch DDP does the same): parameters and
            # buffers (BatchNorm running statistics, num_batches_tracked) come from rank 0 of the group
            src = dist.get_global_rank(group, 0) if group is not None else 0
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, src=src, group=group)
        device = params[0].device
        cap = int(bucket_mb * (1 << 20) / 4)
        self.buckets: List[_Bucket] = []
        cur, size = [], 0
        for p in reversed(params):
            cur.append(p)
            size += p.numel()
            if size >= cap:
                self.buckets.append(_Bucket(cur, device))
                cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, device))
        self._where = {}
        self._handles = []
        self._slotted = []
        for b in self.buckets:
            for i, p in enumerate(b.params):
                self._where[p] = (b, i)
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
                if zero_copy and p.is_cuda and p.dim() == 4:
                    from .kernels import register_grad_slot
                    register_grad_slot(p, b.flat[b.offsets[i]:b.offsets[i] + p.numel()])
                    self._slotted.append(p)

    def _hook(self, p):
        b, i = self._where[p]
        # weight gradients may still be in flight on the side stream (kernels.wgrad_stream): pack and reduce there,
        # ordered after both streams, so the main stream never waits for them during backward
        side = _wgrad_side_stream() if p.grad.is_cuda else None
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            side.wait_event(ev)
            ctx = torch.cuda.stream(side)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            if p.grad.data_ptr() != b.flat.data_ptr() + 4 * b.offsets[i]:
                b.view(i).copy_(p.grad)          # (gradients written straight into their slot need no packing)
            b.pending -= 1
            assert b.pending >= 0, "two backward passes without GradReducer.finish() in between"
            if b.pending == 0:
                op = dist.ReduceOp.AVG if self.avg_native else dist.ReduceOp.SUM
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket and store the averaged gradients into param.grad."""
        if self.buckets[0].flat.is_cuda:
            # gradients may have been packed on the weight-gradient side stream: what follows on the current stream
            # (the fallback reductions, the copy-back) is ordered after it, whether or not backward's own join ran
            from .kernels import wgrad_side_streams
            for side in wgrad_side_streams():
                torch.cuda.current_stream().wait_stream(side)
        for b in self.buckets:
            if b.pending != 0:
                # parameters that received no gradient this step (unused branch): reduce what we have
                for i, p in enumerate(b.params):
                    if p.grad is None:
                        b.view(i).zero_()
                op = dist.ReduceOp.AVG if self.avg_native else dist.ReduceOp.SUM
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        for b in self.buckets:
            b.work.wait()
            if not self.avg_native:
                b.flat.div_(self.world)
            for i, p in enumerate(b.params):
                if p.grad is None:
                    p.grad = b.view(i).clone()
                elif p.grad.data_ptr() != b.flat.data_ptr() + 4 * b.offsets[i]:
                    p.grad.copy_(b.view(i))      # (a gradient that lives in its slot was averaged in place)
            b.pending = len(b.params)
            b.work = None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        if self._slotted:
            from .kernels import unregister_grad_slots
            unregister_grad_slots(self._slotted)
            self._slotted = []

    def param_checksum(self) -> float:
        """sum over ranks of |local checksum - rank-0 checksum| of all parameters: 0.0 iff the replicas are in sync"""
        with torch.no_grad():
            local = torch.stack([p.detach().double().sum() for b in self.buckets for p in b.params]).sum().reshape(1)
            ref = local.clone()
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            dist.broadcast(ref, src=src, group=self.group)
            spread = (local - ref).abs()
            dist.all_reduce(spread, group=self.group)
        return float(spread)
