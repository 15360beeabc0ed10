"""AdamW for the train step, one CUDA launch for all parameters (``grb_adamw_step``).

Same update rule, arguments and ``state_dict`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq``) as
``torch.optim.AdamW``, which is what the reference instantiates through its Hydra optimizer
partial (configs/model/hstu.yaml, models/generative_recommenders.py:254-322 configure_optimizers);
``amsgrad`` / ``maximize`` / sparse gradients are not part of that configuration and raise.

Why not ``torch.optim.AdamW(fused=True)``: at C2 the 131 263 x 256 fp32 item table alone is 941 MB
of traffic per step; ATen's multi-tensor kernel moves it at ~3.6 TB/s, this one is written for the
HBM roofline.  CUDA fp32 parameters only — there is no CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, amsgrad: bool = False,
                 maximize: bool = False):
        if amsgrad or maximize:
            raise NotImplementedError("FusedAdamW: amsgrad / maximize are not implemented")
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("FusedAdamW: lr, eps and weight_decay must be non-negative")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"FusedAdamW: invalid betas {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      amsgrad=False, maximize=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        fn = _lib.lib().grb_adamw_step
        for group in self.param_groups:
            lr = group["lr"]
            lr = float(lr.item()) if isinstance(lr, torch.Tensor) else float(lr)
            beta1, beta2 = group["betas"]
            # parameters are bucketed by (device, step): normally one bucket per group
            buckets: Dict[Tuple[torch.device, float], List[Tuple[torch.Tensor, torch.Tensor, dict]]] = {}
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise NotImplementedError("FusedAdamW: sparse gradients are not supported")
                if not p.is_cuda or p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise RuntimeError(
                        "FusedAdamW: CUDA float32 parameters and gradients only (no CPU path); got "
                        f"{p.device} {p.dtype} / grad {g.dtype}")
                if not p.is_contiguous():
                    raise NotImplementedError("FusedAdamW: parameters must be contiguous")
                st = self.state[p]
                if not st:
                    st["step"] = 0.0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                step = st["step"]
                step = float(step.item()) if isinstance(step, torch.Tensor) else float(step)
                st["step"] = step + 1.0
                if not g.is_contiguous():
                    g = g.contiguous()
                buckets.setdefault((p.device, step + 1.0), []).append((p, g, st))
            for (device, step), items in buckets.items():
                n = len(items)
                arr = C.c_void_p * n
                ps = arr(*[p.data_ptr() for p, _, _ in items])
                gs = arr(*[g.data_ptr() for _, g, _ in items])
                ms = arr(*[s["exp_avg"].data_ptr() for _, _, s in items])
                vs = arr(*[s["exp_avg_sq"].data_ptr() for _, _, s in items])
                ns = (C.c_int64 * n)(*[p.numel() for p, _, _ in items])
                # (switching the device costs milliseconds per step in a multi-process job — measured:
                #  2-GPU end-to-end 84 k -> 38 k sequences/s — so only do it when it is needed)
                ctx = (contextlib.nullcontext() if device.index == torch.cuda.current_device()
                       else torch.cuda.device(device))
                with ctx, _lib.timed("adamw_step"):
                    _lib.check(fn(n, ps, gs, ms, vs, ns, lr, beta1, beta2, group["eps"],
                                  group["weight_decay"], 1.0 - beta1 ** step, 1.0 - beta2 ** step,
                                  _lib.stream_ptr(device)))
        return loss
