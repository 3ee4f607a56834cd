"""Drop-in for the hot-path names of the reference's ``run_nerf_helpers.py``:

    img2mse, mse2psnr, to8b, Embedder, get_embedder, NeRF, ndc_rays, sample_pdf, raw2outputs

Same names, argument meaning and error behaviour; the arithmetic runs in the sm_100a kernels of
``libdlnerf_b200.so`` (no CPU fallback).  Reference lines are cited per function.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .ops import _ptr
from .plan import NetShape, Plan, build_plan

Tensor = torch.Tensor

# Misc (run_nerf_helpers.py:19-21)
# run_nerf_helpers.py:16 is torch.mean((x - y) ** 2): the same arithmetic through ONE autograd node (mse_loss: 2 kernels
# forward, 1 backward instead of 3 + 5) -- on the drop-in route the host time autograd spends on the loss's Sub / Pow /
# Mean backward nodes (~0.4 ms for the three losses) sits between the forward's last kernel and the first backward
# kernel, i.e. it is GPU idle time
img2mse = lambda x, y: torch.nn.functional.mse_loss(x, y)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


# --------------------------------------------------------------------------------------------------
# Positional encoding (run_nerf_helpers.py:25-73)
# --------------------------------------------------------------------------------------------------
class Embedder:
    """Same constructor kwargs as the reference (:26-52).  Only the configuration get_embedder builds
    (include_input, log-sampled octaves, [sin, cos]) has a kernel; anything else raises."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.create_embedding_fn()

    def create_embedding_fn(self):
        kw = self.kwargs
        if not (kw.get("include_input", True) and kw.get("log_sampling", True) and kw.get("input_dims", 3) == 3
                and kw.get("num_freqs") == kw.get("max_freq_log2") + 1):
            raise NotImplementedError("only the get_embedder configuration is implemented on the B200 path")
        self.n_freqs = int(kw["num_freqs"])
        self.out_dim = 3 + 6 * self.n_freqs

    def embed(self, inputs: Tensor) -> Tensor:
        return ops.posenc(inputs, self.n_freqs)


class _IdentityEmbed(nn.Identity):
    n_freqs = 0


def get_embedder(multires, i=0):
    """(:58-73) returns (embed_fn, out_dim); i == -1 -> identity, 3 channels."""
    if i == -1:
        return _IdentityEmbed(), 3
    embedder_obj = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                            log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=embedder_obj):
        return eo.embed(x)

    embed.n_freqs = embedder_obj.n_freqs      # lets the fused renderer recognise the encoding
    return embed, embedder_obj.out_dim


# --------------------------------------------------------------------------------------------------
# Model (run_nerf_helpers.py:77-174)
# --------------------------------------------------------------------------------------------------
_BWD_SIDE = {}          # device -> [side stream, "an MLP backward of the running autograd pass is on it"]


def _overlapped_backward(net: "NeRF", d_out, saved, run):
    """The drop-in route's counterpart of train_step's second stream.  Autograd runs the fine network's MLP backward
    first (it was recorded last) and then the coarse compositing + MLP backward, which do not depend on it: the FIRST MLP
    backward of an autograd pass is launched on a side stream and joined when the pass ends
    (``queue_callback``), so the rest of the pass overlaps its ~1.5 ms of dgrad / wgrad kernels.  Only when the gradients
    are handed over by reference (every ``p.grad`` is None, as after ``optimizer.zero_grad()``): an accumulating
    ``p.grad += g`` would run on the main stream before the join.  ``DLN_NO_BWD_OVERLAP=1`` switches it off."""
    dev = d_out.device
    ent = _BWD_SIDE.get(dev)
    if ent is None:
        ent = _BWD_SIDE[dev] = [torch.cuda.Stream(device=dev), False]
    if (ent[1] or os.environ.get("DLN_NO_BWD_OVERLAP") or torch.cuda.is_current_stream_capturing()
            or any(p.grad is not None for p in net._ordered_params())):
        return run()
    side, main = ent[0], torch.cuda.current_stream(dev)
    side.wait_stream(main)
    for t in [d_out] + [x for x in saved if torch.is_tensor(x)]:
        t.record_stream(side)               # allocated on the main stream, read on the side stream
    with torch.cuda.stream(side):
        grads = run()
    ent[1] = True

    def join():
        main.wait_stream(side)
        ent[1] = False
    torch.autograd.Variable._execution_engine.queue_callback(join)
    return grads


class _MLP(torch.autograd.Function):
    """Fused MLP forward + (dgrad chain, wgrad GEMMs) backward.  Gradients flow to the parameters
    only: on the reference's training path the network inputs never require grad."""

    @staticmethod
    def forward(ctx, net: "NeRF", mode, a, b, P, keep, *params):
        # `keep` is decided by the caller, where the grad mode is still visible: inside forward() grad mode is
        # off and needs_input_grad reports requires_grad of the parameters even under torch.no_grad()
        out, saved = net._run_forward(mode, a, b, P, keep=keep)
        ctx.net, ctx.saved, ctx.P = net, saved, P
        return out

    @staticmethod
    def backward(ctx, d_out):
        net: "NeRF" = ctx.net
        saved, P = ctx.saved, ctx.P
        grads = _overlapped_backward(net, d_out, saved, lambda: net._run_backward(d_out, saved, P))
        ctx.saved = None
        return (None, None, None, None, None, None) + tuple(grads)


class _MLPSem(torch.autograd.Function):
    """``_MLP`` for a network with the semantic head: returns (raw4 [P,4], sem [G,K], point_logits [P,K] | empty).
    ``sem`` are the logits summed over groups of ``S`` consecutive points (run_nerf_helpers.py:589; S = 1: per
    point); ``point_logits`` only fills the semantic columns of a returned ``raw`` and carries no gradient."""

    @staticmethod
    def forward(ctx, net: "NeRF", mode, a, b, P, S, want_points, keep, *params):
        out, saved, sem, pts = net._run_forward(mode, a, b, P, keep=keep, sem_group=S, sem_points=want_points)
        ctx.net, ctx.saved, ctx.P = net, saved, P
        if pts is None:
            pts = out.new_empty(0)
        ctx.mark_non_differentiable(pts)
        if sem is None:
            sem = out.new_empty(0)
            ctx.mark_non_differentiable(sem)
        return out, sem, pts

    @staticmethod
    def backward(ctx, d_out, d_sem, _d_pts):
        net: "NeRF" = ctx.net
        saved, P = ctx.saved, ctx.P
        if torch.is_tensor(d_sem):
            d_sem.record_stream(_BWD_SIDE.setdefault(d_out.device, [torch.cuda.Stream(device=d_out.device), False])[0])
        grads = _overlapped_backward(net, d_out, saved, lambda: net._run_backward(d_out, saved, P, d_sem=d_sem))
        ctx.saved = None
        return (None,) * 8 + tuple(grads)


class NeRF(nn.Module):
    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=False,
                 semantic_num_classes=None):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        self.semantic_num_classes = semantic_num_classes
        # identical registration order / names / shapes / init to the reference (:90-105)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W)
                                        for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        if semantic_num_classes:        # (:107-111) two Linear layers, no activation; registered last
            self.semantic_linear = nn.Sequential(nn.Linear(W, W // 2), nn.Linear(W // 2, semantic_num_classes))
        else:
            self.semantic_linear = False
        self.output_ch = output_ch
        self._shape = NetShape(D=D, W=W, input_ch=input_ch, input_ch_views=input_ch_views, output_ch=output_ch,
                               skips=tuple(skips), use_viewdirs=use_viewdirs,
                               semantic_num_classes=int(semantic_num_classes or 0))
        self._plan: Optional[Plan] = None
        self._dev_state = None

    # ------------------------------------------------------------------ device-side state
    def _ordered_params(self) -> List[nn.Parameter]:
        # called several times per training step: resolve the names once (load_state_dict copies in place and
        # .to()/.cuda() swap .data, so the Parameter objects stay the same)
        cache = self.__dict__.get("_ordered_cache")
        if cache is None:
            d = dict(self.named_parameters())
            cache = [d[name] for name, _ in self._shape.param_shapes()]
            self.__dict__["_ordered_cache"] = cache
        return cache

    def _state(self):
        """Flat fp32 parameter buffer (each nn.Parameter is re-pointed to a view of it), packed bf16
        weight blobs, device copies of the pack jobs / wgrad items."""
        if self._plan is None:
            self._plan = build_plan(self._shape)
        pl = self._plan
        params = self._ordered_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("dlnerf_b200.NeRF: parameters are on %s; the B200 path has no CPU fallback" % dev)
        st = self._dev_state
        names = self.__dict__.get("_param_names")
        if names is None:
            names = self.__dict__["_param_names"] = [n for n, _ in self._shape.param_shapes()]
        # every parameter must still be the view of the flat buffer it was made (one list comparison: the per-parameter
        # Python loop cost ~40 us per call, 4 calls per drop-in step)
        stale = st is None or st["device"] != dev or [p.data_ptr() for p in params] != st["ptrs"]
        if stale:
            flat = torch.zeros(pl.n_flat, device=dev, dtype=torch.float32)     # parameters (+ folded operands M, b')
            for p, n in zip(params, names):
                view = flat[pl.offsets[n]: pl.offsets[n] + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view

            def upload(items):
                raw = b"".join(bytes(i) for i in items)
                return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)

            st = dict(device=dev, flat=flat, wf=torch.zeros(pl.fwd_blob_bytes, device=dev, dtype=torch.uint8),
                      wb=torch.zeros(pl.bwd_blob_bytes, device=dev, dtype=torch.uint8),
                      fwd_jobs=upload(pl.fwd_jobs), bwd_jobs=upload(pl.bwd_jobs), items=upload(pl.wgrad),
                      version=None, sms=torch.cuda.get_device_properties(dev).multi_processor_count)
            st["ptrs"] = [p.data_ptr() for p in params]
            self._dev_state = st
        return st

    def _pack(self, st, force=False):
        params = self._ordered_params()
        ver = tuple(p._version for p in params)
        if st["version"] == ver and not force:
            return
        pl = self._plan
        lib, s = L.lib(), ops._stream()
        if pl.fold:
            L.call("dln_mlp_fold", st["flat"].data_ptr(), *self._fold_args(), s, tag="fold_feature")
        if pl.sem is not None:
            L.call("dln_sem_fold", st["flat"].data_ptr(), C.byref(pl.sem), s, tag="fold_semantic")
        L.call("dln_mlp_pack_weights", st["flat"].data_ptr(), st["fwd_jobs"].data_ptr(), len(pl.fwd_jobs),
                                         st["wf"].data_ptr(), s, tag="pack_weights(fwd)")
        L.call("dln_mlp_pack_weights", st["flat"].data_ptr(), st["bwd_jobs"].data_ptr(), len(pl.bwd_jobs),
                                         st["wb"].data_ptr(), s, tag="pack_weights(bwd)")
        st["version"] = ver

    def _fold_args(self):
        pl, O = self._plan, self._plan.offsets
        return (O["views_linears.0.weight"], self.W + self.input_ch_views, O["feature_linear.weight"],
                O["feature_linear.bias"], O["views_linears.0.bias"], pl.off_M, pl.off_bM)

    # ------------------------------------------------------------------ kernels
    @property
    def sem_K(self) -> int:
        """Semantic logits appended to every output row (0: no head, or a head the reference never evaluates)."""
        return self._shape.sem_K

    def _run_forward(self, mode, a, b, P, keep, force_pack=False, sem_group=0, sem_points=False, sms=None):
        """One forward chain launch.  Returns (out, saved) -- and, when the semantic head is asked for through
        ``sem_group`` (points per summed group, run_nerf_helpers.py:589) / ``sem_points`` (per-point logits for
        ``raw``), (out, saved, sem [G,K] | None, point_logits [P,K] | None).  The head reads the kept activations,
        so an inference pass with semantics also writes the stash."""
        st = self._state()
        self._pack(st, force_pack)
        pl = self._plan
        dev = st["device"]
        n_tiles = (P + L.TILE_ROWS - 1) // L.TILE_ROWS
        out = torch.empty(P, self._shape.out_ch, device=dev, dtype=torch.float32)
        args = L.ChainArgs()
        args.P = P
        strat = None
        if isinstance(mode, tuple):              # ("rays", strat): the kernel computes the stratified depths itself
            mode, strat = mode
        if mode == "rays":                       # a = ray_batch [N, C], b = z_vals [N, S]
            args.rays, args.ray_stride, args.vd_col = a.data_ptr(), a.stride(0), a.shape[1] - 3
            args.S = b.shape[1]
            if strat is None:
                args.z = b.data_ptr()
            else:
                # fused stratified sampling (run_nerf.py:571-593): b is an OUTPUT, filled by the tile prologues
                args.z_gen, args.z_lindisp = b.data_ptr(), int(bool(strat.get("lindisp")))
                if strat.get("rng") is not None:
                    args.z_rng_state, args.z_rng_offset = strat["rng"][0].ptr(), int(strat["rng"][1])
        else:                                    # a = x [P, in]
            args.x, args.x_ld = a.data_ptr(), a.stride(0)
        args.wblob, args.fblob, args.out = st["wf"].data_ptr(), st["flat"].data_ptr(), out.data_ptr()
        saved = None
        want_sem = pl.sem is not None and (sem_group > 0 or sem_points)
        if (sem_group > 0 or sem_points) and pl.sem is None:
            raise RuntimeError("semantic logits requested from a NeRF built without semantic_num_classes / view directions")
        stash = None
        if keep or want_sem:
            stash = torch.empty(n_tiles * pl.fwd_slots * L.SLAB_BYTES, device=dev, dtype=torch.uint8)
            args.stash = stash.data_ptr()
        if keep:
            masks = torch.empty(pl.mask_slots * n_tiles * 2 * 128 * 4, device=dev, dtype=torch.int32)
            args.masks = masks.data_ptr()
            saved = (stash, masks)
        L.call("dln_mlp_chain", C.byref(pl.fwd), C.byref(args), min(st["sms"], sms) if sms else st["sms"], ops._stream(),
               tag="mlp_fwd D=%d" % self.D)
        if not want_sem:
            return out, saved
        K = pl.sem.K
        sem = pts = hsum = None
        if sem_group > 0:
            G = (P + sem_group - 1) // sem_group
            sem = torch.empty(G, K, device=dev, dtype=torch.float32)
            hsum = torch.empty(G, self.W, device=dev, dtype=torch.float32) if keep else None
            L.call("dln_sem_head_fwd", stash.data_ptr(), pl.fwd_slots, pl.h_last_slot, P, int(sem_group),
                   st["flat"].data_ptr(), C.byref(pl.sem), _ptr(hsum), sem.data_ptr(), K, ops._stream(),
                   tag="sem_head_fwd")
        if sem_points:
            pts = torch.empty(P, K, device=dev, dtype=torch.float32)
            L.call("dln_sem_head_fwd", stash.data_ptr(), pl.fwd_slots, pl.h_last_slot, P, 1, st["flat"].data_ptr(),
                   C.byref(pl.sem), None, pts.data_ptr(), K, ops._stream(), tag="sem_point_logits")
        if keep:
            saved = (stash, masks, hsum, int(sem_group))
        return out, saved, sem, pts

    def _run_backward(self, d_out, saved, P, gflat=None, sms=None, d_sem=None, wgrad_sms=None):
        """dgrad chain + wgrad.  ``gflat`` (flat fp32 [n_flat]) accumulates across calls when given (ray-chunked
        steps); otherwise a zeroed buffer is allocated.  ``d_sem`` [G, K]: gradient of the summed semantic logits
        (``sem`` of ``_run_forward``); its input gradient enters the dgrad chain as one fp32 row per group."""
        st = self._state()
        pl = self._plan
        dev = st["device"]
        stash_f, masks = saved[0], saved[1]
        if gflat is None:
            gflat = torch.zeros(pl.n_flat, device=dev, dtype=torch.float32)
        elif pl.n_flat > pl.n_params:
            gflat[pl.n_params:].zero_()        # dM / db' / dSw / dsc scratch of THIS call (the buffer accumulates across calls)
        sem_G = None
        if d_sem is not None and pl.sem is not None and len(saved) > 2 and saved[2] is not None:
            hsum, S = saved[2], saved[3]
            ds = d_sem.reshape(hsum.shape[0], pl.sem.K)
            ds = ds.contiguous() if ds.dtype == torch.float32 else ds.float().contiguous()
            sem_G = torch.empty_like(hsum)
            L.call("dln_sem_head_bwd", ds.data_ptr(), pl.sem.K, hsum.data_ptr(), P, S, st["flat"].data_ptr(),
                   gflat.data_ptr(), C.byref(pl.sem), sem_G.data_ptr(), ops._stream(), tag="sem_head_bwd")
        n_tiles = (P + L.TILE_ROWS - 1) // L.TILE_ROWS
        d = d_out.reshape(P, self._shape.out_ch)
        d = d.contiguous() if d.dtype == torch.float32 else d.float().contiguous()
        stash_b = torch.empty(n_tiles * pl.bwd_slots * L.SLAB_BYTES, device=dev, dtype=torch.uint8)
        args = L.ChainArgs()
        args.P = P
        args.wblob, args.fblob = st["wb"].data_ptr(), st["flat"].data_ptr()
        args.d_out, args.stash, args.masks = d.data_ptr(), stash_b.data_ptr(), masks.data_ptr()
        if sem_G is not None:
            args.sem_g, args.sem_g_div = sem_G.data_ptr(), S
        lib, s = L.lib(), ops._stream()
        # `sms` caps the persistent dgrad grid so that a concurrent kernel on another stream keeps some SMs
        L.call("dln_mlp_chain", C.byref(pl.bwd), C.byref(args), min(st["sms"], sms) if sms else st["sms"], s,
               tag="mlp_dgrad D=%d" % self.D)
        n_items = len(pl.wgrad)
        # ONE wave (n_items * splits <= #SMs): every item's CTA k then walks the same tile range at about the same
        # pace, so slabs two items share (dZ of the skip layer, the last hidden layer, dZ_views) are found in L2 by
        # the second reader and neighbouring slots of a tile are read together; two waves or per-item split counts
        # proportional to the bytes were both measured 10-50 % slower.
        splits = int(max(1, min(n_tiles, (min(st["sms"], wgrad_sms) if wgrad_sms else st["sms"]) // n_items)))
        # deterministic reduction (default): per-CTA shares in a scratch, added in split order by a second kernel;
        # DLN_WGRAD_ATOMICS=1 selects fp32 atomics straight into the gradient buffer instead
        partial = None
        if not os.environ.get("DLN_WGRAD_ATOMICS"):
            partial = torch.empty(n_items * splits * L.WGRAD_PARTIAL_FLOATS, device=dev, dtype=torch.float32)
        L.call("dln_mlp_wgrad", st["items"].data_ptr(), n_items, splits, stash_f.data_ptr(), pl.fwd_slots,
               stash_b.data_ptr(), pl.bwd_slots, n_tiles, gflat.data_ptr(), _ptr(partial), s, tag="mlp_wgrad D=%d" % self.D)
        if pl.fold:
            L.call("dln_mlp_unfold_grads", st["flat"].data_ptr(), gflat.data_ptr(), *self._fold_args(), s,
                   tag="unfold_grads")
        if sem_G is not None:
            L.call("dln_sem_unfold_grads", st["flat"].data_ptr(), gflat.data_ptr(), C.byref(pl.sem), s,
                   tag="unfold_semantic")
        grads = []
        shapes = self.__dict__.get("_param_shapes")
        if shapes is None:
            shapes = self.__dict__["_param_shapes"] = self._shape.param_shapes()
        for (name, shp), p in zip(shapes, self._ordered_params()):
            o = pl.offsets[name]
            grads.append(gflat[o: o + p.numel()].view(p.shape) if p.requires_grad else None)
        return grads

    def _keep(self) -> bool:
        """Whether a forward pass must write the activation stash / ReLU masks: only when autograd will come back
        for the parameter gradients."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self._ordered_params())

    # ------------------------------------------------------------------ public interface
    @ops.on_device_of(1)
    def forward(self, x: Tensor) -> Tensor:
        """(:113-145) x[..., input_ch + input_ch_views] -> [..., 4] (view dirs) or [..., output_ch]."""
        n_in = self.input_ch + (self.input_ch_views if self.use_viewdirs else 0)
        if x.shape[-1] < n_in:
            raise RuntimeError("expected at least %d input channels, got %d" % (n_in, x.shape[-1]))
        xs = ops._f32(x, "NeRF.forward").reshape(-1, x.shape[-1])
        P = xs.shape[0]
        if self.sem_K:                  # [rgb, alpha, semantic logits] (:139-140); S = 1: per-point logits
            out, sem, _ = _MLPSem.apply(self, "x", xs, None, P, 1, False, self._keep(), *self._ordered_params())
            out = torch.cat([out, sem], -1)
        else:
            out = _MLP.apply(self, "x", xs, None, P, self._keep(), *self._ordered_params())
        return out.reshape(*x.shape[:-1], out.shape[-1])

    @staticmethod
    def fused_sampling_available() -> bool:
        """The CTA-pair chain kernel (default) can compute the stratified depths in its tile prologue."""
        return os.environ.get("DLN_CHAIN", "2")[:1] != "1"

    @ops.on_device_of(1)
    def forward_rays(self, ray_batch: Tensor, z_vals: Tensor, semantic: bool = False, point_logits: bool = False,
                     strat: Optional[dict] = None):
        """Fused path of run_nerf.py:595 + run_network (:60-74) + forward: points o + d*z are formed,
        encoded (positions per sample, the unit view direction once per ray) and pushed through the MLP
        inside one kernel.  ray_batch is the packed [N, 8|11] batch of render(); returns raw[N, S, 4].

        With the semantic head, ``semantic`` / ``point_logits`` make it return (raw[N, S, 4], sem_preds[N, K] | None,
        logits[N, S, K] | None): the per-ray logits of raw2outputs (helpers:589: the UNWEIGHTED sum of the per-sample
        logits; differentiable) and the per-sample logits that fill raw[..., 4:] in the reference (values only).

        ``strat`` = dict(rng=(RngState, offset) | None, lindisp=bool): fused stratified sampling (run_nerf.py:571-593) --
        ``z_vals`` is then an uninitialised [N, S] tensor the kernel FILLS (same bits as ``ops.stratified_z``) while it
        encodes the points; the coarse pass of render_rays / train_step uses it."""
        rb, z = ops._f32(ray_batch, "forward_rays"), ops._f32(z_vals, "forward_rays")
        if self.use_viewdirs and rb.shape[1] < 11:
            raise RuntimeError("use_viewdirs=True needs the unit view direction in the last 3 ray columns")
        N, S = z.shape
        if (semantic or point_logits) and not self.sem_K:
            raise RuntimeError("this NeRF has no semantic head (semantic_num_classes / use_viewdirs)")
        mode = "rays" if strat is None else ("rays", strat)
        if semantic or point_logits:
            out, sem, pts = _MLPSem.apply(self, mode, rb, z, N * S, S if semantic else 0, bool(point_logits),
                                          self._keep(), *self._ordered_params())
            return (out.reshape(N, S, 4), sem if semantic else None,
                    pts.reshape(N, S, -1) if point_logits else None)
        out = _MLP.apply(self, mode, rb, z, N * S, self._keep(), *self._ordered_params())
        return out.reshape(N, S, out.shape[-1])

    def load_weights_from_keras(self, weights):
        """(:147-174) same index arithmetic as the reference: transposed Keras kernels."""
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        with torch.no_grad():
            for i in range(self.D):
                self.pts_linears[i].weight.copy_(torch.from_numpy(np.transpose(weights[2 * i])))
                self.pts_linears[i].bias.copy_(torch.from_numpy(np.transpose(weights[2 * i + 1])))
            k = 2 * self.D
            self.feature_linear.weight.copy_(torch.from_numpy(np.transpose(weights[k])))
            self.feature_linear.bias.copy_(torch.from_numpy(np.transpose(weights[k + 1])))
            self.views_linears[0].weight.copy_(torch.from_numpy(np.transpose(weights[k + 2])))
            self.views_linears[0].bias.copy_(torch.from_numpy(np.transpose(weights[k + 3])))
            self.rgb_linear.weight.copy_(torch.from_numpy(np.transpose(weights[k + 4])))
            self.rgb_linear.bias.copy_(torch.from_numpy(np.transpose(weights[k + 5])))
            self.alpha_linear.weight.copy_(torch.from_numpy(np.transpose(weights[k + 6])))
            self.alpha_linear.bias.copy_(torch.from_numpy(np.transpose(weights[k + 7])))


# --------------------------------------------------------------------------------------------------
# Ray helpers (run_nerf_helpers.py:320-337)
# --------------------------------------------------------------------------------------------------
def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Near-plane shift + projective warp (run_nerf_helpers.py:320-337) with torch ops; render() uses the
    one-launch `ops.pack_rays` for device tensors, this function keeps the reference's name and signature."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx, sy = -1. / (W / (2. * focal)), -1. / (H / (2. * focal))
    # the reference's operation order, (sx * o_x) / o_z, so the results are bit-identical to it (:326-333)
    o = torch.stack([sx * rays_o[..., 0] / rays_o[..., 2], sy * rays_o[..., 1] / rays_o[..., 2],
                     1. + 2. * near / rays_o[..., 2]], -1)
    d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2]),
                     sy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2]),
                     -2. * near / rays_o[..., 2]], -1)
    return o, d


def get_rays_np(H, W, focal, c2w):
    """run_nerf_helpers.py:285-300 on the device: (rays_o, rays_d), each [H, W, 3] fp32 CUDA tensors (a stack of poses
    [n, 3, 4] gives [n, H, W, 3] in one launch -- the list comprehension of run_nerf.py:1126).  Values are bit-identical
    to the numpy function; call ``.cpu().numpy()`` on them where an array is really needed."""
    return ops.gen_rays(int(H), int(W), float(focal), c2w)


def get_rays_by_coord_np(H, W, focal, c2w, coords):
    """run_nerf_helpers.py:303-318 on the device: rays through the fractional pixel positions coords[N, 2] = (x, y) of the
    LiDAR / COLMAP depth points, in the coordinates' dtype (float64 coordinates -> float64 rays, as numpy promotes)."""
    return ops.gen_rays_by_coord(int(H), int(W), float(focal), c2w, coords)


def get_rays_cropped_feature_loss_new(H, W, focal, c2w, nH=32, nW=32, gradH=2, gradW=2, device=None, _start=None,
                                      _perm=None):
    """run_nerf_helpers.py:430-494: a random nH x nW crop, its pixels randomly split into gradH*gradW rays rendered with
    gradients and the rest without.  Returns the reference's three lists
    ``[grad_rays_o, grad_rays_d, grad_points], [no_grad_rays_o, no_grad_rays_d, no_grad_points], [start_w, end_w,
    start_h, end_h]`` with every tensor on the device.  Randomness as in the reference: the crop corner from
    ``np.random.randint`` (:436-437, same two draws in the same order), the split from ``torch.randperm`` (:466, drawn on
    the device); ``_start=(start_w, start_h)`` / ``_perm`` inject them (parity tests)."""
    H, W = int(H), int(W)
    num_w, num_h = W - nW + 1, H - nH + 1
    if _start is None:
        start_w = int(np.random.randint(0, num_w))
        start_h = int(np.random.randint(0, num_h))
    else:
        start_w, start_h = int(_start[0]), int(_start[1])
    end_w, end_h = start_w + nW - 1, start_h + nH - 1
    device = ops._cuda_device(device, c2w)
    perm = torch.randperm(nH * nW, device=device) if _perm is None else torch.as_tensor(_perm).to(device)
    o, d, pts = ops.gen_rays_patch(H, W, float(focal), c2w, start_w, start_h, nH, nW, perm)
    k = gradH * gradW
    return [o[:k], d[:k], pts[:k]], [o[k:], d[k:], pts[k:]], [start_w, end_w, start_h, end_h]


# --------------------------------------------------------------------------------------------------
# Hierarchical sampling (run_nerf_helpers.py:497-540)
# --------------------------------------------------------------------------------------------------
@ops.on_device_of(0)
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """Same signature as the reference.  Unless det, the uniforms u[..., N_samples] (:509) are drawn inside the
    kernel (Philox, seeded from torch's seed) instead of by torch.rand.  The result carries no gradient (the reference
    detaches it at run_nerf.py:634 before any use)."""
    lead = list(bins.shape[:-1])
    u, rng = None, None
    if not det:
        st = ops.default_rng(bins.device, 2)
        rng = (st, st.next_offsets(1))
    if pytest:
        np.random.seed(0)
        u = None if det else torch.tensor(np.random.rand(*(lead + [N_samples])), dtype=torch.float32,
                                         device=bins.device)
    return ops.sample_pdf(bins.detach(), weights.detach(), N_samples, u, rng=rng)


# --------------------------------------------------------------------------------------------------
# Compositing (run_nerf_helpers.py:542-595)
# --------------------------------------------------------------------------------------------------
@ops.on_device_of(0)
def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, semantic_loss=False):
    """Returns (rgb_map, disp_map, acc_map, weights, depth_map[, semantic_class_preds]), differentiable w.r.t.
    raw.  semantic_loss: the per-ray logits are the UNWEIGHTED sum of raw[..., 4:] over the samples (:586-593)."""
    noise, rng = None, None
    if raw_noise_std > 0.:
        st = ops.default_rng(raw.device, 2)       # N(0,1) drawn in-kernel (:565), regenerated by the backward
        rng = (st, st.next_offsets(1))
        if pytest:       # the reference's hook draws UNIFORM numbers here (:567-571)
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)), dtype=torch.float32, device=raw.device)
    maps = ops.composite(raw, z_vals, rays_d, noise, float(raw_noise_std), bool(white_bkgd), rng=rng)
    if semantic_loss:
        return tuple(maps) + (ops.sample_sum(raw, 4),)
    return maps
