"""Adam on the flat parameter buffers.

``FlatAdam`` is a ``torch.optim.Optimizer`` whose state has torch.optim.Adam's layout (per parameter ``step``,
``exp_avg``, ``exp_avg_sq``; same ``param_groups`` keys), so the reference's checkpoint code
(``optimizer.state_dict()`` / ``load_state_dict``, run_nerf.py:448-477, :1872-1883) and its learning-rate decay
(``param_group['lr'] = new_lrate``, :1843-1847) work unchanged -- but ``step()`` is ONE kernel per network
(``dln_adam_step`` over the flat fp32 buffer every ``NeRF`` keeps its parameters in) followed by the bf16 weight
re-pack, instead of ~10 element-wise launches per parameter tensor (48 tensors for the coarse + fine pair).

    optimizer = dn.FlatAdam([model, model_fine], lr=args.lrate, betas=(0.9, 0.999))     # replaces run_nerf.py:440
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch

from . import _lib as L
from . import ops
from .run_nerf_helpers import NeRF


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, nets: Sequence[NeRF], lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        nets = [n for n in nets if n is not None]
        if not nets or not all(isinstance(n, NeRF) for n in nets):
            raise TypeError("FlatAdam takes the dlnerf_b200.NeRF modules whose parameters it updates")
        self._nets: List[NeRF] = list(nets)
        params = [p for n in self._nets for p in n._ordered_params()]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False))
        self._flat_state = None

    # ------------------------------------------------------------------ flat moments, exposed per parameter
    def _ensure_state(self):
        if self._flat_state is not None and all(fs["flat"] is n._state()["flat"] for fs, n in zip(self._flat_state, self._nets)):
            return
        flat_state = []
        for n in self._nets:
            st = n._state()
            pl = n._plan
            m = torch.zeros_like(st["flat"])
            v = torch.zeros_like(st["flat"])
            names = [nm for nm, _ in n._shape.param_shapes()]
            for p, nm in zip(n._ordered_params(), names):
                o = pl.offsets[nm]
                old = self.state.get(p, {})
                mv, vv = m[o:o + p.numel()].view(p.shape), v[o:o + p.numel()].view(p.shape)
                if "exp_avg" in old:          # state loaded from a checkpoint (or a previous flat buffer)
                    mv.copy_(old["exp_avg"])
                    vv.copy_(old["exp_avg_sq"])
                step = old.get("step", torch.tensor(0.0))
                self.state[p] = {"step": torch.as_tensor(float(step)), "exp_avg": mv, "exp_avg_sq": vv}
            flat_state.append(dict(flat=st["flat"], m=m, v=v))
        self._flat_state = flat_state

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat_state = None            # moments are re-flattened (copied into the flat buffers) on the next step

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._ensure_state()
        group = self.param_groups[0]
        lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
        for n, fs in zip(self._nets, self._flat_state):
            params = n._ordered_params()
            if any(p.grad is None for p in params if p.requires_grad):
                continue                   # torch.optim skips parameters without a gradient; here: the whole net
            pl = n._plan
            names = [nm for nm, _ in n._shape.param_shapes()]
            g0 = params[0].grad
            base = g0.data_ptr() - 4 * pl.offsets[names[0]]
            contiguous = g0.dtype == torch.float32 and base % 16 == 0 and all(
                p.grad.is_contiguous() and p.grad.data_ptr() == base + 4 * pl.offsets[nm] for p, nm in zip(params, names))
            if contiguous:
                gptr, keep = base, None    # the gradients already are views of one flat buffer in parameter order
            else:
                keep = torch.zeros_like(fs["flat"])
                for p, nm in zip(params, names):
                    keep[pl.offsets[nm]:pl.offsets[nm] + p.numel()].copy_(p.grad.reshape(-1))
                gptr = keep.data_ptr()
            t = int(self.state[params[0]]["step"]) + 1
            L.call("dln_adam_step", fs["flat"].data_ptr(), gptr, fs["m"].data_ptr(), fs["v"].data_ptr(),
                   pl.n_params, lr, float(b1), float(b2), eps, t, float(grad_scale), ops._stream(), tag="adam_step")
            for p in params:
                self.state[p]["step"] = torch.as_tensor(float(t))
            n._pack(n._state(), force=True)        # bf16 weight stages follow the fp32 master copy
        return loss
