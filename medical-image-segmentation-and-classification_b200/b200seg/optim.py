"""FusedClipAdamW — clip_grad_norm_(max_norm) + AdamW(lr, betas, eps, weight_decay) for all parameters in two kernel
launches of libb200seg.so (b2_grad_sqnorm_multi, b2_adamw_multi).

Drop-in for the optimizer half of the reference's hot loop (utils/helpers.py:251,333-335):

    opt = FusedClipAdamW(model.parameters(), lr=lr, weight_decay=5e-4, max_norm=1.0)
    loss.backward(); opt.step()                     # == clip_grad_norm_(params, 1.0); AdamW.step()

It is a torch.optim.Optimizer (param_groups / state_dict / load_state_dict / LR schedulers work; the learning rate is
mirrored into a device scalar so a captured CUDA graph sees scheduler updates).  The moments live in buffers whose
addresses are baked into the device-side pointer table, so load_state_dict() copies the loaded moments INTO those
buffers (and restores the step counter, kept as state[p]["step"] like torch.optim.AdamW) instead of swapping tensors.  Gradients may be re-allocated every step
(zero_grad(set_to_none=True)): the device-side pointer table is refreshed from a pinned staging buffer when any
pointer changed; inside a CUDA-graph capture the upload is captured too, from a private snapshot of the table.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import kernels as K
from .kernels import _p, _stream

_CHUNK = 1 << 16          # elements per block


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdamW supports a single parameter group (as the reference's train() uses)")
        self.max_norm = float(max_norm)
        self._params = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not self._params:
            raise ValueError("no trainable parameters")
        dev = self._params[0].device
        for p in self._params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise ValueError("FusedClipAdamW needs fp32 CUDA parameters")
            st = self.state[p]
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] = torch.zeros((), dtype=torch.float32)        # host mirror, refreshed by state_dict()
        # block -> (tensor, chunk) map
        bt, bc = [], []
        for i, p in enumerate(self._params):
            for c in range((p.numel() + _CHUNK - 1) // _CHUNK):
                bt.append(i)
                bc.append(c)
        self._nblocks = len(bt)
        self._block_tensor = torch.tensor(bt, dtype=torch.int32, device=dev)
        self._block_chunk = torch.tensor(bc, dtype=torch.int32, device=dev)
        n = len(self._params)
        self._refs_host = torch.zeros((n, 5), dtype=torch.int64).pin_memory()
        self._refs_dev = torch.zeros((n, 5), dtype=torch.int64, device=dev)
        self._grad_ptrs = [0] * n
        self._captured_tables = []
        self._sqnorm = torch.zeros((), dtype=torch.float64, device=dev)
        self._step = torch.zeros((), dtype=torch.float32, device=dev)    # device-side step counter (bias correction)
        self._lr = torch.full((), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self.total_norm = torch.zeros((), dtype=torch.float32, device=dev)
        for i, p in enumerate(self._params):
            st = self.state[p]
            self._refs_host[i, 0] = p.data_ptr()
            self._refs_host[i, 2] = st["exp_avg"].data_ptr()
            self._refs_host[i, 3] = st["exp_avg_sq"].data_ptr()
            self._refs_host[i, 4] = p.numel()
        # re-pack table (b2_pack_ref rows: 10 x int64) for the cached bf16 copies of the conv weights
        self._pack_serial = -1
        self._pack_dev = None
        self._pack_n = 0
        self._pack_items = 0
        self._pack_keep = []
        self._pack_hosts = []

    def _moment_buffers(self):
        return [(self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) for p in self._params]

    def state_dict(self):
        """torch.optim.Optimizer.state_dict() with the per-parameter `step` entries filled from the device counter"""
        step = float(self._step)
        for p in self._params:
            self.state[p]["step"] = torch.tensor(step, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """The kernels read the moments through a table of raw addresses captured at construction (and possibly baked
        into a CUDA graph), so the loaded moments are copied into the EXISTING buffers; the step counter is restored
        from the saved per-parameter `step`."""
        keep = self._moment_buffers()
        super().load_state_dict(state_dict)
        params_now = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if len(params_now) != len(self._params):
            raise ValueError("loaded optimizer state does not match the parameter list")
        steps = set()
        for p, (m, v) in zip(self._params, keep):
            st = self.state[p]
            if "exp_avg" in st:
                if st["exp_avg"].shape != m.shape:
                    raise ValueError("loaded optimizer state has a different parameter shape")
                if st["exp_avg"].data_ptr() != m.data_ptr():
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                steps.add(float(st.get("step", 0.0)))
            else:                       # a state dict saved before the first step
                m.zero_()
                v.zero_()
                steps.add(0.0)
            st["exp_avg"], st["exp_avg_sq"] = m, v
        if len(steps) > 1:
            raise ValueError(f"FusedClipAdamW keeps one step counter; the loaded state has {sorted(steps)}")
        self._step.fill_(steps.pop() if steps else 0.0)
        g = self.param_groups[0]
        self._lr_host = float(g["lr"])
        self._lr.fill_(self._lr_host)

    def _grad_for(self, p):
        g = p.grad
        if g is None:
            raise RuntimeError("FusedClipAdamW.step(): a trainable parameter has no gradient")
        if g.stride() != p.stride():      # element-wise kernels need identical memory order
            g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
            p.grad = g
        return g

    def _refresh_refs(self):
        changed = False
        for i, p in enumerate(self._params):
            ptr = self._grad_for(p).data_ptr()
            if ptr != self._grad_ptrs[i]:
                self._grad_ptrs[i] = ptr
                self._refs_host[i, 1] = ptr
                changed = True
        if changed:
            if torch.cuda.is_current_stream_capturing():
                # becomes a captured H2D copy: replays re-upload the static capture-time addresses.  The source must
                # never change afterwards, so it is a private pinned snapshot kept alive with the optimizer.
                snap = self._refs_host.clone().pin_memory()
                self._captured_tables.append(snap)
                self._refs_dev.copy_(snap, non_blocking=True)
                self._grad_ptrs = [0] * len(self._params)      # force a fresh upload on the next eager step
            else:
                self._refs_dev.copy_(self._refs_host, non_blocking=True)

    def sync_lr(self):
        """mirror param_groups[0]['lr'] into the device scalar the kernels read (call outside graph capture / replay)"""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self._lr_host = lr
            self._lr.fill_(lr)

    def _refresh_pack_table(self):
        """device table of every cached packed copy (kernels.packed) of the parameters this optimizer updates"""
        if self._pack_serial == K.pack_serial():
            return
        packs = K.packs_of(self._params)
        rows, start = [], 0
        for pk in packs:
            tco, tci = (pk.cout + 31) // 32, (pk.cin + 31) // 32
            assert pk.code == 2 or pk.ksize <= 3
            items = tco * ((pk.cols + 31) // 32) if pk.code == 2 else tco * tci      # one item = a 32 x 32 tile, all taps
            lo = (pk.cout & 0xFFFFFFFF) | (pk.cin << 32)
            hi = (pk.ksize & 0xFFFFFFFF) | (pk.code << 32)
            rows.append([pk.ptr, pk.wf.data_ptr(), pk.wd.data_ptr() if pk.wd is not None else 0, lo, hi,
                         pk.strides[0], pk.strides[1], pk.strides[2], pk.strides[3],
                         (start & 0xFFFFFFFF) | (pk.cols << 32)])                               # one b2_pack_ref (80 B)
            start += items
        self._pack_n, self._pack_items = len(rows), start
        self._pack_keep = packs
        if rows:
            host = torch.tensor(rows, dtype=torch.int64).pin_memory()
            self._pack_hosts.append(host)          # a captured graph may replay this upload: never freed
            dev = self._refs_dev.device
            if len(rows) > 4096:
                raise RuntimeError("FusedClipAdamW: more than 4096 packed weight copies")
            if self._pack_dev is None:      # fixed size: its address is baked into captured graphs
                self._pack_dev = torch.zeros((4096, 10), dtype=torch.int64, device=dev)
            self._pack_dev[:len(rows)].copy_(host, non_blocking=True)
        self._pack_serial = K.pack_serial()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise ValueError("closures are not supported")
        g = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        self._refresh_refs()
        self._refresh_pack_table()
        b1, b2 = g["betas"]
        refs = C.c_void_p(self._refs_dev.data_ptr())
        _lib.call("b2_grad_sqnorm_multi", refs, _p(self._block_tensor), _p(self._block_chunk), self._nblocks, _CHUNK,
                  _p(self._sqnorm), _p(self._step), _stream())
        _lib.call("b2_adamw_multi", refs, _p(self._block_tensor), _p(self._block_chunk), self._nblocks, _CHUNK,
                  _p(self._sqnorm), self.max_norm, _p(self._lr), float(b1), float(b2), float(g["eps"]),
                  float(g["weight_decay"]), _p(self._step), _p(self.total_norm), _stream())
        if self._pack_n:
            # the fp32 masters just changed through raw pointers: refresh every bf16 MMA-layout copy in ONE launch
            _lib.call("b2_pack_weights_multi", C.c_void_p(self._pack_dev.data_ptr()), self._pack_n, self._pack_items,
                      _stream())
        K.bump_param_epoch()
        return None
