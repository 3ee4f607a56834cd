"""torch-facing wrappers over the C ABI (include/dlnerf_b200.h).

PyTorch is used for device memory, streams and autograd plumbing only; every arithmetic stage is one
of the library's sm_100a kernels.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L

Tensor = torch.Tensor


def _stream() -> int:
    """Raw handle of the current stream of the current device.  (torch.cuda.current_stream() builds a Stream object and
    resolves the device index in Python: ~19 us per call, 12 calls per drop-in step; the C accessor takes ~1 us.)"""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def on_device_of(argpos: int = 0, kw: Optional[str] = None):
    """Decorator for the public entry points: run the body with the CUDA device of a tensor argument current, so
    that kernel launches, `_stream()` and the per-device kernel attributes all refer to the tensors' device even
    when the caller's current device is another GPU (one process driving several GPUs)."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            t = kwargs.get(kw) if kw is not None and kw in kwargs else (args[argpos] if len(args) > argpos else None)
            if isinstance(t, (tuple, list)) and t:
                t = t[0]
            if isinstance(t, torch.nn.Module):
                t = next(t.parameters(), None)
            if torch.is_tensor(t) and t.is_cuda and t.device.index != torch.cuda.current_device():
                with torch.cuda.device(t.device):
                    return fn(*args, **kwargs)
            return fn(*args, **kwargs)
        return wrapped
    return deco


def _f32(t: Tensor, what: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError("dlnerf_b200.%s: tensor is on %s; the B200 path has no CPU fallback" % (what, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# In-kernel random numbers (include/dlnerf_b200.h, "In-kernel random numbers")
# ------------------------------------------------------------------------------------------------
class RngState:
    """Device-resident {seed, base} pair of the Philox generator the sampling / compositing kernels carry.  A random
    tensor is named by (seed, base + offset): ``offset`` is passed by value with each launch, ``base`` lives on the
    device (``advance`` bumps it in-stream, so a captured CUDA graph draws fresh numbers on every replay)."""

    def __init__(self, device, seed: Optional[int] = None):
        self.device = torch.device(device)
        self.seed = int(torch.initial_seed() if seed is None else seed) & ((1 << 63) - 1)
        self.state = torch.tensor([self.seed, 0], dtype=torch.int64, device=self.device)
        self.calls = 0                       # host-side tensor counter (drop-in route: offsets instead of `advance`)

    def ptr(self) -> int:
        return self.state.data_ptr()

    def advance(self, inc: int) -> None:
        L.call("dln_rng_advance", self.ptr(), int(inc), _stream(), tag="rng_advance")

    def next_offsets(self, n: int) -> int:
        """Reserve n tensor names; returns the first offset (host-side bookkeeping, no launch)."""
        o = self.calls
        self.calls += n
        return o

    def fill(self, offset: int, kind: str, rows: int, row_len: int) -> Tensor:
        """The draws a kernel generates for tensor (base + offset), as a tensor (tests): kind 'u' uniform row-major,
        'n' normal row-major, 'u_resample' uniform in the slot order of the <=64+64 importance_resample path."""
        out = torch.empty(rows, row_len, device=self.device, dtype=torch.float32)
        L.call("dln_rng_fill", self.ptr(), int(offset), {"u": 0, "n": 1, "u_resample": 2}[kind], out.data_ptr(),
               rows, row_len, _stream(), tag="rng_fill")
        return out


Rng = Optional[Tuple[RngState, int]]         # (state, by-value offset) naming one random tensor
_DEFAULT_RNG = {}


def default_rng(device, stream_id: int = 0) -> RngState:
    """Per-device generator state seeded from torch's (``torch.manual_seed`` re-seeds it; ranks of a distributed job
    get distinct streams).  ``stream_id`` separates independent consumers (0: render_rays, 1: train_step)."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    rank = 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank = dist.get_rank()
    except Exception:
        rank = 0
    key = (device, stream_id)
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03 + stream_id * 0x94D049BB133111EB) \
        & ((1 << 63) - 1)
    st = _DEFAULT_RNG.get(key)
    if st is None or st.seed != seed:
        st = _DEFAULT_RNG[key] = RngState(device, seed)
    return st


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


def stratified_z(ray_batch: Tensor, n_samples: int, t_rand: Optional[Tensor] = None, lindisp: bool = False,
                 rng: Rng = None) -> Tensor:
    """run_nerf.py:571-593.  ray_batch[N, >=8] with near/far at columns 6/7.  Jitter: ``t_rand`` (a tensor of U[0,1)
    draws) or ``rng`` (drawn in-kernel) or neither (bin centres)."""
    rb = _f32(ray_batch, "stratified_z")
    N = rb.shape[0]
    z = torch.empty(N, n_samples, device=rb.device, dtype=torch.float32)
    if rng is not None and t_rand is None:
        L.call("dln_stratified_z_rng", rb.data_ptr(), rb.stride(0), rng[0].ptr(), int(rng[1]), z.data_ptr(), N,
               n_samples, int(bool(lindisp)), _stream(), tag="stratified_z")
        return z
    tr = None if t_rand is None else _f32(t_rand, "stratified_z")
    if tr is not None and tuple(tr.shape) != (N, n_samples):
        raise ValueError("t_rand must be [N, N_samples]")
    L.call("dln_stratified_z", rb.data_ptr(), rb.stride(0), _ptr(tr), z.data_ptr(), N, n_samples,
                                     int(bool(lindisp)), _stream(), tag="stratified_z")
    return z


def _cuda_device(device, like=None) -> torch.device:
    """An indexed CUDA device: `device` if given, else the device of the CUDA tensor `like`, else the current one."""
    if device is not None:
        dev = torch.device(device)
    elif torch.is_tensor(like) and like.is_cuda:
        dev = like.device
    else:
        dev = torch.device("cuda")
    if dev.type != "cuda":
        raise RuntimeError("dlnerf_b200: ray generation runs on the GPU; there is no CPU fallback")
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


def _pose34(c2w, dev, dtype) -> Tensor:
    """[..., >=3, >=4] camera-to-world matrices -> contiguous [..., 3, 4] of `dtype` on `dev`."""
    t = c2w if torch.is_tensor(c2w) else torch.as_tensor(np.asarray(c2w))
    return t[..., :3, :4].to(device=dev, dtype=dtype).contiguous()


def gen_rays(H: int, W: int, focal: float, c2w, out: Optional[Tensor] = None, device=None):
    """get_rays_np (run_nerf_helpers.py:285-300) for one pose [3,4] or a stack [n,3,4] in ONE launch.  Returns
    (rays_o, rays_d) of shape [(n,) H, W, 3]; with ``out`` = a [n*H*W, 3, 3] bank the rays are written straight into its
    rows 0 (origin) and 1 (direction) (run_nerf.py:1126-1147 without the host round trip)."""
    dev = _cuda_device(device, c2w)
    with torch.cuda.device(dev):
        p = _pose34(c2w, dev, torch.float32)
        single = p.dim() == 2
        p = p.reshape(-1, 3, 4)
        n = p.shape[0]
        if out is None:
            o = torch.empty(n, H, W, 3, device=dev, dtype=torch.float32)
            d = torch.empty_like(o)
            po, pd, stride = o.data_ptr(), d.data_ptr(), 3
        else:
            if tuple(out.shape) != (n * H * W, 3, 3) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
                raise ValueError("out must be a contiguous fp32 [n*H*W, 3, 3] bank on the poses' device")
            o, d = out[:, 0], out[:, 1]
            po, pd, stride = out.data_ptr(), out.data_ptr() + 12, 9
        L.call("dln_gen_rays", p.data_ptr(), n, int(H), int(W), float(focal), po, pd, stride, _stream(), tag="gen_rays")
    if out is None and single:
        return o[0], d[0]
    return o, d


def gen_rays_by_coord(H: int, W: int, focal: float, c2w, coords, device=None):
    """get_rays_by_coord_np (run_nerf_helpers.py:303-318): coords[N,2] fractional (x, y) pixels -> (rays_o, rays_d)
    [N,3] in the coordinates' floating dtype (numpy's promotion: float64 coordinates give float64 rays)."""
    ct = coords if torch.is_tensor(coords) else torch.as_tensor(np.asarray(coords))
    dev = _cuda_device(device, ct)
    dtype = torch.float64 if ct.dtype == torch.float64 else torch.float32
    with torch.cuda.device(dev):
        ct = ct.to(device=dev, dtype=dtype).reshape(-1, 2).contiguous()
        p = _pose34(c2w, dev, dtype)
        if p.dim() != 2:
            raise ValueError("get_rays_by_coord takes one [3,4] pose")
        N = ct.shape[0]
        o = torch.empty(N, 3, device=dev, dtype=dtype)
        d = torch.empty_like(o)
        if N == 0:
            return o, d
        L.call("dln_gen_rays_by_coord", p.data_ptr(), ct.data_ptr(), N, int(H), int(W), float(focal),
               int(dtype == torch.float64), o.data_ptr(), d.data_ptr(), 3, _stream(), tag="gen_rays_by_coord")
    return o, d


def gen_rays_patch(H: int, W: int, focal: float, c2w, start_w: int, start_h: int, nH: int, nW: int, perm: Tensor):
    """The permuted crop of get_rays_cropped_feature_loss_new (run_nerf_helpers.py:430-494): rays and (row, col)
    crop positions of the pixels perm[0], perm[1], ... (flat row-major indices into the nH x nW crop)."""
    dev = perm.device
    if not perm.is_cuda:
        raise ValueError("perm must live on the GPU")
    with torch.cuda.device(dev):
        p = _pose34(c2w, dev, torch.float32)
        pm = perm.to(torch.int64).contiguous()
        n = pm.numel()
        o = torch.empty(n, 3, device=dev, dtype=torch.float32)
        d = torch.empty_like(o)
        pts = torch.empty(n, 2, device=dev, dtype=torch.int64)
        L.call("dln_gen_rays_patch", p.data_ptr(), int(H), int(W), float(focal), int(start_w), int(start_h), int(nH),
               int(nW), pm.data_ptr(), n, o.data_ptr(), d.data_ptr(), pts.data_ptr(), _stream(), tag="gen_rays_patch")
    return o, d, pts


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
    def forward(ctx, raw, z_vals, rays_d, noise, noise_std, white_bkgd, rng):
        raw_c, z_c, d_c = _f32(raw, "raw2outputs"), _f32(z_vals, "raw2outputs"), _f32(rays_d, "raw2outputs")
        N, S, Cc = raw_c.shape
        nz = None if noise is None else _f32(noise, "raw2outputs")
        if nz is not None:
            rng = None
        dev = raw_c.device
        rgb = torch.empty(N, 3, device=dev)
        disp, acc, depth = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
        w = torch.empty(N, S, device=dev)
        if rng is not None:
            L.call("dln_composite_fwd_rng", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), rng[0].ptr(),
                   int(rng[1]), float(noise_std), int(white_bkgd), rgb.data_ptr(), disp.data_ptr(), acc.data_ptr(),
                   w.data_ptr(), depth.data_ptr(), N, S, _stream(), tag="composite_fwd")
        else:
            L.call("dln_composite_fwd", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), _ptr(nz),
                   float(noise_std), int(white_bkgd), rgb.data_ptr(), disp.data_ptr(),
                   acc.data_ptr(), w.data_ptr(), depth.data_ptr(), N, S, _stream(), tag="composite_fwd")
        ctx.save_for_backward(raw_c, z_c, d_c, nz if nz is not None else torch.empty(0, device=dev))
        ctx.cfg = (float(noise_std), int(white_bkgd), nz is not None)
        ctx.rng = rng
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_w, g_depth):
        raw_c, z_c, d_c, nz = ctx.saved_tensors
        noise_std, white, has_noise = ctx.cfg
        N, S, Cc = raw_c.shape
        gs = [None if g is None else _f32(g, "raw2outputs.backward") for g in (g_rgb, g_disp, g_acc, g_w, g_depth)]
        d_raw = torch.empty_like(raw_c)
        if ctx.rng is not None:          # the same (seed, base + offset) as the forward: the draws are regenerated
            L.call("dln_composite_bwd_rng", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), ctx.rng[0].ptr(),
                   int(ctx.rng[1]), noise_std, white, _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]), _ptr(gs[4]),
                   d_raw.data_ptr(), N, S, _stream(), tag="composite_bwd")
        else:
            L.call("dln_composite_bwd", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(),
                   nz.data_ptr() if has_noise else None, noise_std, white,
                   _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]), _ptr(gs[4]),
                   d_raw.data_ptr(), N, S, _stream(), tag="composite_bwd")
        return d_raw, None, None, None, None, None, None


def composite(raw: Tensor, z_vals: Tensor, rays_d: Tensor, noise: Optional[Tensor], noise_std: float,
              white_bkgd: bool, rng: Rng = None):
    """(rgb_map, disp_map, acc_map, weights, depth_map); `noise` are UNSCALED N(0,1) draws or None; with ``rng`` (and
    no ``noise``) the draws are generated in-kernel -- the backward regenerates them, so the device-side ``base`` of
    the state must not be advanced between the two (the drop-in route names its tensors by offset only)."""
    if raw.dim() != 3 or raw.shape[-1] < 4:
        raise ValueError("raw must be [N, S, >=4]")
    if raw.shape[1] > 256:
        raise NotImplementedError("composite kernels handle up to 256 samples per ray")
    return _Composite.apply(raw, z_vals, rays_d, noise, noise_std, white_bkgd, rng if noise_std > 0. else None)


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
                             depth_mode: int, depth_norm: float, loss_sums: Tensor, rng: Rng = None,
                             coefs_dev: Optional[Tensor] = None) -> Tensor:
    """north_star part 5: d raw with the RGB-MSE / LiDAR-depth loss gradient formed in-kernel.  ``coefs_dev`` (device
    fp32 [coef_rgb, coef_depth, depth_norm]) replaces the three by-value scalars; ``rng`` draws the density noise
    in-kernel (the same tensor name as the forward compositing of this pass)."""
    raw_c, z_c, d_c = _f32(raw, "fused_loss"), _f32(z_vals, "fused_loss"), _f32(rays_d, "fused_loss")
    N, S, Cc = raw_c.shape
    nz = None if noise is None else _f32(noise, "fused_loss")
    d_raw = torch.empty_like(raw_c)
    if nz is not None or not noise_std > 0.:
        rng = None
    if rng is not None or coefs_dev is not None:
        if coefs_dev is None:
            coefs_dev = torch.tensor([coef_rgb, coef_depth, depth_norm], dtype=torch.float32).to(raw_c.device)
        L.call("dln_composite_bwd_fused_loss_dev", raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), _ptr(nz),
               rng[0].ptr() if rng is not None else None, int(rng[1]) if rng is not None else 0, float(noise_std),
               int(white_bkgd), _ptr(target_rgb), _ptr(target_depth), _ptr(ray_weights), int(n_rgb),
               coefs_dev.data_ptr(), int(depth_mode), loss_sums.data_ptr(), d_raw.data_ptr(), N, S, _stream(),
               tag="composite_bwd_fused_loss")
        return d_raw
    L.call("dln_composite_bwd_fused_loss", 
        raw_c.data_ptr(), Cc, z_c.data_ptr(), d_c.data_ptr(), _ptr(nz), float(noise_std), int(white_bkgd),
        _ptr(target_rgb), _ptr(target_depth), _ptr(ray_weights), int(n_rgb), float(coef_rgb), float(coef_depth),
        int(depth_mode), float(depth_norm), loss_sums.data_ptr(), d_raw.data_ptr(), N, S, _stream(), tag="composite_bwd_fused_loss")
    return d_raw


# ------------------------------------------------------------------------------------------------
def sample_pdf(bins: Tensor, weights: Tensor, n_samples: int, u: Optional[Tensor] = None,
               return_debug: bool = False, rng: Rng = None):
    """sample_pdf, run_nerf_helpers.py:497-540.  u=None -> deterministic linspace (det=True) unless ``rng`` (the
    uniforms are then drawn in-kernel)."""
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
    if rng is not None and uu is None:
        L.call("dln_sample_pdf_rng", b2.data_ptr(), B, 0, w2.data_ptr(), B - 1, B, rng[0].ptr(), int(rng[1]), n_samples,
               out.data_ptr(), None, 0, None, _ptr(cdf), _ptr(inds), N, _stream(), tag="sample_pdf")
    else:
        L.call("dln_sample_pdf", b2.data_ptr(), B, 0, w2.data_ptr(), B - 1, B, _ptr(uu), n_samples, out.data_ptr(),
               None, 0, None, _ptr(cdf), _ptr(inds), N, _stream(), tag="sample_pdf")
    out = out.reshape(*lead, n_samples)
    if return_debug:
        return out, cdf.reshape(*lead, B), inds.reshape(*lead, n_samples)
    return out


def importance_resample(z_vals: Tensor, weights: Tensor, n_importance: int, u: Optional[Tensor] = None,
                        return_debug: bool = False, rng: Rng = None):
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
    if rng is not None and uu is None:
        L.call("dln_sample_pdf_rng", z.data_ptr(), S, 1, w.data_ptr() + 4, S, S - 1, rng[0].ptr(), int(rng[1]),
               n_importance, zs.data_ptr(), z.data_ptr(), S, zm.data_ptr(), _ptr(cdf), _ptr(inds), N, _stream(),
               tag="importance_resample")
    else:
        L.call("dln_sample_pdf", z.data_ptr(), S, 1, w.data_ptr() + 4, S, S - 1, _ptr(uu), n_importance,
               zs.data_ptr(), z.data_ptr(), S, zm.data_ptr(), _ptr(cdf), _ptr(inds), N, _stream(),
               tag="importance_resample")
    if return_debug:
        return zs, zm, cdf, inds
    return zs, zm


def composite_resample(raw: Tensor, z_vals: Tensor, rays_d: Tensor, noise: Optional[Tensor], noise_std: float,
                       white_bkgd: bool, n_importance: int, u: Optional[Tensor] = None, rng_noise: Rng = None,
                       rng_u: Rng = None):
    """``composite`` (values only, no autograd node) + ``importance_resample`` of its weights in one launch
    (run_nerf.py:600-636) for 64 coarse and <= 64 new samples; the same bits as the two calls.  Returns
    (rgb_map, disp_map, acc_map, weights, depth_map, z_samples, z_merged)."""
    raw_c, z, d = _f32(raw, "composite_resample"), _f32(z_vals, "composite_resample"), _f32(rays_d, "composite_resample")
    N, S, Cc = raw_c.shape
    if S != 64 or not 1 <= n_importance <= 64:
        raise ValueError("composite_resample: 64 coarse and at most 64 new samples (got %d + %d)" % (S, n_importance))
    nz = None if noise is None else _f32(noise, "composite_resample")
    uu = None if u is None else _f32(u, "composite_resample")
    rn = None if nz is not None else rng_noise
    ru = None if uu is not None else rng_u
    dev = raw_c.device
    rgb = torch.empty(N, 3, device=dev)
    disp, acc, depth = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
    w = torch.empty(N, S, device=dev)
    zs = torch.empty(N, n_importance, device=dev)
    zm = torch.empty(N, S + n_importance, device=dev)
    L.call("dln_composite_resample_fwd", raw_c.data_ptr(), Cc, z.data_ptr(), d.data_ptr(), _ptr(nz),
           rn[0].ptr() if rn else None, int(rn[1]) if rn else 0, float(noise_std), int(white_bkgd), rgb.data_ptr(),
           disp.data_ptr(), acc.data_ptr(), w.data_ptr(), depth.data_ptr(), _ptr(uu), ru[0].ptr() if ru else None,
           int(ru[1]) if ru else 0, int(n_importance), zs.data_ptr(), zm.data_ptr(), N, S, _stream(),
           tag="composite_resample")
    return rgb, disp, acc, w, depth, zs, zm


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
