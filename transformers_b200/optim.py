"""Optimizer step behind the reference's own hooks.

``B200AdamW`` is a ``torch.optim.Optimizer`` with ``torch.optim.AdamW``'s constructor, param groups and state layout
(``step`` / ``exp_avg`` / ``exp_avg_sq``), so the reference's ``Trainer`` takes it unchanged::

    Trainer(model, args, optimizer_cls_and_kwargs=(B200AdamW, {"lr": 2e-5, "betas": (0.9, 0.999), "eps": 1e-8}))   # trainer.py:385,1200-1203
    Trainer(model, args, optimizers=(B200AdamW(model.parameters(), lr=2e-5), None))

``clip_grad_norm_`` mirrors ``torch.nn.utils.clip_grad_norm_`` (what ``Trainer._clip_grad_norm`` reaches through
accelerate, trainer.py:2538-2542).  With ``max_grad_norm`` given to the optimizer the clip coefficient never leaves the
device and is applied inside the AdamW kernel (no separate scaling pass over the gradients, no host sync); set
``TrainingArguments.max_grad_norm=0`` then.

Each param group is ONE kernel launch over all its tensors (multi-tensor table, include/b200_ops.h); bf16 parameters and
gradients, moments in the parameter dtype like torch (or fp32 with ``state_dtype=torch.float32``); ``master_weights=True``
keeps an fp32 master copy of every parameter (``state["master"]``) that the update runs on -- bf16 parameters alone lose
every update below half an ulp."""
from __future__ import annotations

import math

import torch

from . import ops
from ._lib import B200Error


def _tables(entries, device):
    """entries: list of (param_ptr, grad_ptr, m_ptr, v_ptr, numel) -> (table int64 [n,6], chunk map int32 [c,2]) on device."""
    chunk = ops.optim_chunk_elems()
    rows, cmap = [], []
    for i, e in enumerate(entries):
        p, g, m, v, n = e[:5]
        rows.append((p, g, m, v, n, e[5] if len(e) > 5 else 0))
        cmap.extend((i, c) for c in range((n + chunk - 1) // chunk))
    table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 6)
    cm = torch.tensor(cmap, dtype=torch.int32).reshape(-1, 2)
    if device.type == "cuda":
        return table.pin_memory().to(device, non_blocking=True), cm.pin_memory().to(device, non_blocking=True)
    return table.to(device), cm.to(device)


def _check_param(p):
    from . import modules as M

    if not M._on_b200(p):
        raise B200Error(f"B200AdamW: parameters must live on the B200 (got {p.device}); there is no CPU fallback")
    if p.dtype not in _DTYPES or not p.is_contiguous():
        raise B200Error(f"B200AdamW: parameters must be contiguous bfloat16 tensors, got {p.dtype} contiguous={p.is_contiguous()}")
    if p.grad.dtype != p.dtype or not p.grad.is_contiguous() or p.grad.is_sparse:
        raise B200Error("B200AdamW: gradients must be dense, contiguous and of the parameter dtype")


_DTYPES = (torch.bfloat16,)


def clip_grad_norm_(parameters, max_norm: float, norm_type: float = 2.0) -> torch.Tensor:
    """Drop-in for ``torch.nn.utils.clip_grad_norm_`` (L2 only): scales the gradients in place by
    min(1, max_norm / (total_norm + 1e-6)) and returns the total norm (0-dim fp32 tensor on the device, no host sync)."""
    if float(norm_type) != 2.0:
        raise B200Error("clip_grad_norm_: only the L2 norm is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.zeros(())
    for p in ps:
        _check_param(p)
    table, cmap = _tables([(0, p.grad.data_ptr(), 0, 0, p.grad.numel()) for p in ps], ps[0].device)
    out = ops.grad_norm(table, cmap, max_norm)
    ops.grad_scale_(table, cmap, out[1:2])
    _bump_versions([p.grad for p in ps])
    return out[0]


def _bump_versions(tensors) -> None:
    """The kernels write through raw pointers, which autograd's version counters do not see: bump them so that everything
    keyed on ``_version`` (modules.fused_weight's cached q|k|v and gate|up concatenations, saved-tensor checks) notices the
    in-place update exactly as it would after ``torch.optim.AdamW.step``."""
    torch._C._increment_version(list(tensors))


class B200AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: float | None = None, state_dtype: torch.dtype | None = None, master_weights: bool = False,
                 **unsupported):
        for k in ("amsgrad", "maximize", "capturable", "differentiable"):
            if unsupported.pop(k, False):
                raise B200Error(f"B200AdamW: {k}=True is not supported")
        for k in ("foreach", "fused"):
            unsupported.pop(k, None)  # implementation selectors of torch.optim.AdamW: meaningless here
        if unsupported:
            raise TypeError(f"B200AdamW: unexpected arguments {sorted(unsupported)}")
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("B200AdamW: invalid hyper-parameters")
        if state_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("B200AdamW: state_dtype must be None (parameter dtype), torch.bfloat16 or torch.float32")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.state_dtype = state_dtype
        # fp32 master copy of every parameter (state["master"]): the update runs on it, the bf16 parameter is its rounding
        self.master_weights = bool(master_weights)
        self.grad_norm = None  # device tensor with the pre-clip global norm of the last step (when max_grad_norm is set)

    def load_state_dict(self, state_dict):
        """``Optimizer.load_state_dict`` casts every floating-point state tensor except ``step`` to the parameter dtype
        (bf16): an fp32 ``master`` copy or fp32 moments would come back as bf16 while the kernel reads them as ``float*``.
        Restore them from the incoming state dict itself (no round trip through bf16) in the dtypes this optimizer uses."""
        from itertools import chain

        super().load_state_dict(state_dict)
        saved_ids = chain.from_iterable(g["params"] for g in state_dict["param_groups"])
        params = chain.from_iterable(g["params"] for g in self.param_groups)
        for sid, p in zip(saved_ids, params):
            src = state_dict["state"].get(sid)
            if src is None or p not in self.state:
                continue
            for key in ("exp_avg", "exp_avg_sq", "master"):
                if torch.is_tensor(src.get(key)):
                    want = torch.float32 if key == "master" else (self.state_dtype or p.dtype)
                    self.state[p][key] = src[key].detach().to(device=p.device, dtype=want, copy=True).contiguous()
            if self.master_weights and "master" not in self.state[p]:  # checkpoint written without master weights
                self.state[p]["master"] = p.detach().to(torch.float32, copy=True)

    def _init_state(self, p):
        st = self.state[p]
        if not st:
            dt = self.state_dtype or p.dtype
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, dtype=dt, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, dtype=dt, memory_format=torch.preserve_format)
            if self.master_weights:
                st["master"] = p.detach().to(torch.float32, copy=True)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            for p in ps:
                _check_param(p)
            if ps:
                groups.append((group, ps))
        if not groups:
            return loss
        device = groups[0][1][0].device
        coef = None
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            every = [p for _, ps in groups for p in ps]
            table, cmap = _tables([(0, p.grad.data_ptr(), 0, 0, p.grad.numel()) for p in every], device)
            out = ops.grad_norm(table, cmap, self.max_grad_norm)
            self.grad_norm, coef = out[0], out[1:2]
        for group, ps in groups:
            launches = {}  # (step, moments are fp32) -> table entries; normally one key per group
            for p in ps:
                st = self._init_state(p)
                st["step"] += 1
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if m.dtype != v.dtype or m.dtype not in (torch.float32, p.dtype):
                    raise B200Error("B200AdamW: moments must both be fp32 or both have the parameter dtype")
                key = (int(st["step"]), m.dtype == torch.float32)
                for t in (m, v):
                    if t.shape != p.shape or t.device != p.device or not t.is_contiguous():
                        raise B200Error("B200AdamW: optimizer state does not match its parameter (shape / device / layout)")
                master = 0
                if self.master_weights:
                    mw = st.get("master")
                    if mw is None or mw.dtype != torch.float32 or mw.shape != p.shape or mw.device != p.device or not mw.is_contiguous():
                        raise B200Error("B200AdamW: state['master'] must be a contiguous fp32 tensor shaped like its parameter")
                    master = mw.data_ptr()
                launches.setdefault(key, []).append((p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), master))
            beta1, beta2 = group["betas"]
            lr = float(group["lr"])  # schedulers write python floats (or 0-dim tensors) into the group
            for (step, fp32), entries in sorted(launches.items()):
                table, cmap = _tables(entries, device)
                ops.adamw_step(table, cmap, state_fp32=fp32, master=self.master_weights, lr=lr, beta1=beta1, beta2=beta2, eps=group["eps"],
                               weight_decay=group["weight_decay"], bias_correction1=1.0 - beta1 ** step,
                               bias_correction2_sqrt=math.sqrt(1.0 - beta2 ** step), grad_scale=coef)
            _bump_versions(ps)
        return loss
