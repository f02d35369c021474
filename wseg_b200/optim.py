"""Drop-in for the reference's ``tool.torchutils.PolyOptimizer`` (tool/torchutils.py:11-33; constructed at
contrast_train.py:91-96, stepped at :397-399) with the parameter update as ONE fused multi-tensor kernel per group
(``wsegk_sgd_step``) instead of a chain of foreach launches — the step after the hot path.

The reference's quirks are part of its behaviour and are reproduced, not fixed:
* ``super().__init__(params, lr, weight_decay)`` passes the weight decay POSITIONALLY into ``torch.optim.SGD``, whose third
  positional argument is ``momentum``: every group runs with momentum = wt_dec (5e-4) and the default weight decay is 0
  (the script then sets ``weight_decay`` per group explicitly);
* the constructor's ``momentum=0.9`` is only used as the POWER of the poly learning-rate schedule;
* the schedule stops updating ``lr`` once ``global_step >= max_step`` (the last value stays).
CUDA fp32 parameters only: there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check


class PolyOptimizer(torch.optim.Optimizer):
    def __init__(self, params, lr, weight_decay, max_step, momentum=0.9):
        defaults = dict(lr=lr, momentum=weight_decay, dampening=0, weight_decay=0, nesterov=False)   # the positional quirk
        super().__init__(params, defaults)
        self.global_step = 0
        self.max_step = max_step
        self.momentum = momentum
        self._initial_lr = [group['lr'] for group in self.param_groups]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.global_step < self.max_step:
            lr_mult = (1 - self.global_step / self.max_step) ** self.momentum
            for i, group in enumerate(self.param_groups):
                group['lr'] = self._initial_lr[i] * lr_mult
        for group in self.param_groups:
            ps, gs, ms, first = [], [], [], None
            for p in group['params']:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("wseg_b200.optim.PolyOptimizer: parameters must be contiguous CUDA fp32 tensors (no CPU path)")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                state = self.state[p]
                new = 'momentum_buffer' not in state or state['momentum_buffer'] is None
                if new:
                    state['momentum_buffer'] = torch.empty_like(p)
                first = new if first is None else first
                if new != first:
                    raise RuntimeError("PolyOptimizer: parameters of one group must enter training together")
                ps.append(p); gs.append(g); ms.append(state['momentum_buffer'])
            if not ps:
                continue
            n = len(ps)
            arr = ctypes.c_void_p * n
            lens = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
            with torch.cuda.device(ps[0].device):
                check(_lib.lib().wsegk_sgd_step(arr(*[p.data_ptr() for p in ps]), arr(*[g.data_ptr() for g in gs]),
                                                arr(*[m.data_ptr() for m in ms]), lens, n, float(group['lr']),
                                                float(group['weight_decay']), float(group['momentum']),
                                                float(group['dampening']), 1 if first else 0,
                                                ctypes.c_void_p(torch.cuda.current_stream(ps[0].device).cuda_stream)),
                      "wsegk_sgd_step")
        self.global_step += 1
        return loss
