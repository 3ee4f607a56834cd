"""torch-facing wrappers over the C ABI (include/dlnerf_b200.h).

PyTorch is used for device memory, streams and autograd plumbing only; every arithmetic stage is one
of the library's sm_100a kernels.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: Tensor, what: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError("dlnerf_b200.%s: tensor is on %s; the B200 path has no CPU fallback" % (what, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
def pack_rays(H: int, W: int, focal: float, rays_o: Tensor, rays_d: Tensor, ndc: bool, near: float, far: float,
              use_viewdirs: bool) -> Tensor:
    """render()'s ray packing in one launch (run_nerf.py:145-183 with ndc_rays, helpers:320-337, near plane 1):
    [..., 3] origins / directions -> ray_batch[N, 8 | 11] = [o, d, near, far, (unit viewdirs of the pre-warp d)]."""
    o = _f32(rays_o, "pack_rays").reshape(-1, 3)
    d = _f32(rays_d, "pack_rays").reshape(-1, 3)
    if o.shape != d.shape:
        raise ValueError("rays_o and rays_d must have the same shape")
    N = o.shape[0]
    out = torch.empty(N, 11 if use_viewdirs else 8, device=o.device, dtype=torch.float32)
    L.call("dln_pack_rays", o.data_ptr(), d.data_ptr(), N, int(bool(ndc)), int(H), int(W), float(focal), 1.0,
           float(near), float(far), int(bool(use_viewdirs)), out.data_ptr(), _stream(), tag="pack_rays")
    return out


def stratified_z(ray_batch: Tensor, n_samples: int, t_rand: Optional[Tensor] = None, lindisp: bool = False) -> Tensor:
    """run_nerf.py:571-593.  ray_batch[N, >=8] with near/far at columns 6/7."""
    rb = _f32(ray_batch, "stratified_z")
    N = rb.shape[0]
    z = torch.empty(N, n_samples, device=rb.device, dtype=torch.float32)
    tr = None if t_rand is None else _f32(t_rand, "stratified_z")
    if tr is not None and tuple(tr.shape) != (N, n_samples):
        raise ValueError("t_rand must be [N, N_samples]")
    L.call("dln_stratified_z", rb.data_ptr(), rb.stride(0), _ptr(tr), z.data_ptr(), N, n_samples,
                                     int(bool(lindisp)), _stream(), tag="stratified_z")
    return z


def posenc(x: Tensor, n_freqs: int) -> Tensor:
    """Embedder.embed, run_nerf_helpers.py:54-55 (forward only, inputs carry no gradient on this path)."""
    xs = _f32(x, "posenc")
    if xs.shape[-1] != 3:
        raise ValueError("posenc expects [..., 3]")
    flat = xs.reshape(-1, 3)
    out = torch.empty(flat.shape[0], 3 + 6 * n_freqs, device=xs.device, dtype=torch.float32)
    L.call("dln_posenc", flat.data_ptr(), out.data_ptr(), flat.shape[0], n_freqs, _stream(), tag="posenc")
    return out.reshape(*xs.shape[:-1], out.shape[-1])


# ------------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    """raw2outputs forward/backward (run_nerf_helpers.py:542-595).  Differentiable w.r.t. ``raw`` only,
    which is all the reference's training graph needs (z_vals / rays_d are leaves without grad and
    z_samples is detached, run_nerf.py:634)."""

    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise, noise_std, white_bkgd):
        raw_c, z_c, d_c = _f32(raw, "raw2outputs"), _f32(z_vals, "raw2outputs"), _f32(rays_d, "raw2outputs")
        N, S, Cc = raw_c.shape
        nz = None if noise is None else _f32(noise, "raw2outputs")
        dev = raw_c.device
        rgb = torch.empty(N, 3, device=dev)
        disp, acc, depth = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
        w = torch.empty(N, S, device=dev)
        L.call("dln_composite_fwd", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), _ptr(nz),
                                          float(noise_std), int(white_bkgd), rgb.data_ptr(), disp.data_ptr(),
                                          acc.data_ptr(), w.data_ptr(), depth.data_ptr(), N, S, _stream(), tag="composite_fwd")
        ctx.save_for_backward(raw_c, z_c, d_c, nz if nz is not None else torch.empty(0, device=dev))
        ctx.cfg = (float(noise_std), int(white_bkgd), nz is not None)
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_w, g_depth):
        raw_c, z_c, d_c, nz = ctx.saved_tensors
        noise_std, white, has_noise = ctx.cfg
        N, S, Cc = raw_c.shape
        gs = [None if g is None else _f32(g, "raw2outputs.backward") for g in (g_rgb, g_disp, g_acc, g_w, g_depth)]
        d_raw = torch.empty_like(raw_c)
        L.call("dln_composite_bwd", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(),
                                          nz.data_ptr() if has_noise else None, noise_std, white,
                                          _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]), _ptr(gs[4]),
                                          d_raw.data_ptr(), N, S, _stream(), tag="composite_bwd")
        return d_raw, None, None, None, None, None


def composite(raw: Tensor, z_vals: Tensor, rays_d: Tensor, noise: Optional[Tensor], noise_std: float,
              white_bkgd: bool):
    """(rgb_map, disp_map, acc_map, weights, depth_map); `noise` are UNSCALED N(0,1) draws or None."""
    if raw.dim() != 3 or raw.shape[-1] < 4:
        raise ValueError("raw must be [N, S, >=4]")
    if raw.shape[1] > 256:
        raise NotImplementedError("composite kernels handle up to 256 samples per ray")
    return _Composite.apply(raw, z_vals, rays_d, noise, noise_std, white_bkgd)


class _SampleSum(torch.autograd.Function):
    """torch.sum(raw[..., c0:], -2) (run_nerf_helpers.py:589) for a caller-supplied raw tensor."""

    @staticmethod
    def forward(ctx, raw, c0):
        raw_c = _f32(raw, "sample_sum")
        N, S, Cc = raw_c.shape
        out = torch.empty(N, Cc - c0, device=raw_c.device)
        L.call("dln_sample_sum", raw_c.data_ptr(), Cc, c0, N, S, out.data_ptr(), _stream(), tag="sample_sum")
        ctx.cfg = (N, S, Cc, c0)
        return out

    @staticmethod
    def backward(ctx, g):
        N, S, Cc, c0 = ctx.cfg
        gc = _f32(g, "sample_sum.backward")
        d_raw = torch.empty(N, S, Cc, device=gc.device)
        L.call("dln_sample_sum_bwd", gc.data_ptr(), Cc, c0, N, S, d_raw.data_ptr(), _stream(), tag="sample_sum_bwd")
        return d_raw, None


def sample_sum(raw: Tensor, c0: int = 4) -> Tensor:
    """Per-ray semantic logits of raw2outputs: the unweighted sum of raw[..., c0:] over the samples."""
    if raw.dim() != 3 or raw.shape[-1] <= c0:
        raise ValueError("raw must be [N, S, >%d] to carry semantic logits" % c0)
    return _SampleSum.apply(raw, c0)


def semantic_ce(sem: Tensor, target: Tensor, n_rgb: int, coef: float, loss_sum: Tensor) -> Tensor:
    """F.cross_entropy(sem[:n_rgb], target) (run_nerf.py:1542, :1546), fused forward + gradient: ADDS the summed
    loss of the first n_rgb rays to loss_sum[0] and returns dsem = coef * (softmax - onehot) (zero rows behind
    n_rgb)."""
    s = _f32(sem, "semantic_ce")
    N, K = s.shape
    t = target.to(device=s.device, dtype=torch.int64).contiguous() if n_rgb > 0 else None
    if t is not None and t.numel() < n_rgb:
        raise ValueError("target_semantic needs one class index per RGB ray")
    dsem = torch.empty_like(s)
    L.call("dln_sem_ce_loss", s.data_ptr(), K, _ptr(t), int(n_rgb), N, K, float(coef), dsem.data_ptr(),
           loss_sum.data_ptr(), _stream(), tag="semantic_ce")
    return dsem


def composite_bwd_fused_loss(raw: Tensor, z_vals: Tensor, rays_d: Tensor, noise: Optional[Tensor], noise_std: float,
                             white_bkgd: bool, target_rgb: Optional[Tensor], target_depth: Optional[Tensor],
                             ray_weights: Optional[Tensor], n_rgb: int, coef_rgb: float, coef_depth: float,
                             depth_mode: int, depth_norm: float, loss_sums: Tensor) -> Tensor:
    """north_star part 5: d raw with the RGB-MSE / LiDAR-depth loss gradient formed in-kernel."""
    raw_c, z_c, d_c = _f32(raw, "fused_loss"), _f32(z_vals, "fused_loss"), _f32(rays_d, "fused_loss")
    N, S, Cc = raw_c.shape
    nz = None if noise is None else _f32(noise, "fused_loss")
    d_raw = torch.empty_like(raw_c)
    L.call("dln_composite_bwd_fused_loss", 
        raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), _ptr(nz), float(noise_std), int(white_bkgd),
        _ptr(target_rgb), _ptr(target_depth), _ptr(ray_weights), int(n_rgb), float(coef_rgb), float(coef_depth),
        int(depth_mode), float(depth_norm), loss_sums.data_ptr(), d_raw.data_ptr(), N, S, _stream(), tag="composite_bwd_fused_loss")
    return d_raw


# ------------------------------------------------------------------------------------------------
def sample_pdf(bins: Tensor, weights: Tensor, n_samples: int, u: Optional[Tensor] = None,
               return_debug: bool = False):
    """sample_pdf, run_nerf_helpers.py:497-540.  u=None -> deterministic linspace (det=True)."""
    b, w = _f32(bins, "sample_pdf"), _f32(weights, "sample_pdf")
    lead = b.shape[:-1]
    B = b.shape[-1]
    if w.shape[-1] != B - 1:
        raise ValueError("weights must have len(bins)-1 entries")
    b2, w2 = b.reshape(-1, B), w.reshape(-1, B - 1)
    N = b2.shape[0]
    uu = None
    if u is not None:
        uu = _f32(u, "sample_pdf").reshape(N, n_samples)
    out = torch.empty(N, n_samples, device=b.device)
    cdf = torch.empty(N, B, device=b.device) if return_debug else None
    inds = torch.empty(N, n_samples, device=b.device, dtype=torch.int64) if return_debug else None
    L.call("dln_sample_pdf", b2.data_ptr(), B, 0, w2.data_ptr(), B - 1, B, _ptr(uu), n_samples, out.data_ptr(),
                                   None, 0, None, _ptr(cdf), _ptr(inds), N, _stream(), tag="sample_pdf")
    out = out.reshape(*lead, n_samples)
    if return_debug:
        return out, cdf.reshape(*lead, B), inds.reshape(*lead, n_samples)
    return out


def importance_resample(z_vals: Tensor, weights: Tensor, n_importance: int, u: Optional[Tensor] = None,
                        return_debug: bool = False):
    """run_nerf.py:632-636 in one kernel: bins = midpoints of z_vals, pdf from weights[..., 1:-1],
    CDF inversion, and the sorted union with z_vals.  Returns (z_samples, z_merged)."""
    z, w = _f32(z_vals, "importance_resample"), _f32(weights, "importance_resample")
    N, S = z.shape
    if S < 3:
        raise ValueError("need at least 3 coarse samples")
    uu = None if u is None else _f32(u, "importance_resample")
    zs = torch.empty(N, n_importance, device=z.device)
    zm = torch.empty(N, S + n_importance, device=z.device)
    cdf = torch.empty(N, S - 1, device=z.device) if return_debug else None
    inds = torch.empty(N, n_importance, device=z.device, dtype=torch.int64) if return_debug else None
    L.call("dln_sample_pdf", z.data_ptr(), S, 1, w.data_ptr() + 4, S, S - 1, _ptr(uu), n_importance,
                                   zs.data_ptr(), z.data_ptr(), S, zm.data_ptr(), _ptr(cdf), _ptr(inds), N, _stream(), tag="importance_resample")
    if return_debug:
        return zs, zm, cdf, inds
    return zs, zm


def searchsorted(a: Tensor, v: Tensor, side: str = "left") -> Tensor:
    """Contract of the vendored torchsearchsorted extension (searchsorted.py:20-53)."""
    if side not in ("left", "right"):
        raise ValueError("side must be 'left' or 'right'")
    aa, vv = _f32(a, "searchsorted"), _f32(v, "searchsorted")
    if aa.dim() != 2 or vv.dim() != 2:
        raise ValueError("input `a` and `v` must be 2-D")
    if not (aa.shape[0] == vv.shape[0] or aa.shape[0] == 1 or vv.shape[0] == 1):
        raise ValueError("`a` and `v` must have the same number of rows or one of them must have only one")
    rows = max(aa.shape[0], vv.shape[0])
    out = torch.empty(rows, vv.shape[1], device=aa.device, dtype=torch.int64)
    L.call("dln_searchsorted", aa.data_ptr(), aa.shape[0], aa.shape[1], vv.data_ptr(), vv.shape[0], vv.shape[1],
                                     out.data_ptr(), int(side == "right"), _stream(), tag="searchsorted")
    return out
