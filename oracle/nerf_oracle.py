"""CPU oracle for the ray-rendering training hot path of mertkiray/depth-lidar-nerf.

TEST INFRASTRUCTURE ONLY.  This file is a plain torch-on-CPU restatement of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product path (``depth-lidar-nerf_b200``) never does and fails loudly when its
CUDA library is missing.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (third-party non-path imports stubbed), runs
it on seeded inputs with the RNG draws injected, checks every function below
against it and commits the vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks the oracle against those vectors
wherever the repo travels (the reference itself does not travel).

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  All functions are dtype-generic: run them in float32 to
mimic the reference, or float64 to get a tight ground truth for tolerances.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# R5  positional encoding            run_nerf_helpers.py:25-73
# --------------------------------------------------------------------------- #
def posenc(x: Tensor, n_freqs: int) -> Tensor:
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].

    run_nerf_helpers.py:41-55 — include_input=True, log_sampling=True, the
    frequency table is 2**linspace(0, L-1, L) and every block is 3 wide."""
    bands = 2.0 ** torch.linspace(0.0, n_freqs - 1, n_freqs, dtype=x.dtype)
    parts = [x]
    for f in bands:
        parts += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(parts, dim=-1)


def posenc_dim(n_freqs: int, i_embed: int = 0) -> int:
    """run_nerf_helpers.py:58-73 (i_embed == -1 -> identity, 3 channels)."""
    return 3 if i_embed == -1 else 3 * (1 + 2 * n_freqs)


# --------------------------------------------------------------------------- #
# R7  NeRF MLP                       run_nerf_helpers.py:77-145
# --------------------------------------------------------------------------- #
@dataclass
class MLPSpec:
    """Static shape of one NeRF network (run_nerf_helpers.py:78-111)."""
    D: int = 8
    W: int = 256
    input_ch: int = 63
    input_ch_views: int = 27
    output_ch: int = 5
    skips: Sequence[int] = (4,)
    use_viewdirs: bool = True
    semantic_num_classes: int = 0      # K > 0: semantic_linear = Linear(W, W/2) -> Linear(W/2, K)  (:107-111)

    def param_shapes(self) -> Dict[str, tuple]:
        """Parameter names/shapes exactly as the reference's state_dict
        (run_nerf_helpers.py:90-111), so checkpoints interoperate."""
        shp: Dict[str, tuple] = {}
        for i in range(self.D):
            if i == 0:
                fan_in = self.input_ch
            elif (i - 1) in self.skips:
                fan_in = self.W + self.input_ch
            else:
                fan_in = self.W
            shp[f"pts_linears.{i}.weight"] = (self.W, fan_in)
            shp[f"pts_linears.{i}.bias"] = (self.W,)
        shp["views_linears.0.weight"] = (self.W // 2, self.input_ch_views + self.W)
        shp["views_linears.0.bias"] = (self.W // 2,)
        if self.use_viewdirs:
            shp["feature_linear.weight"] = (self.W, self.W)
            shp["feature_linear.bias"] = (self.W,)
            shp["alpha_linear.weight"] = (1, self.W)
            shp["alpha_linear.bias"] = (1,)
            shp["rgb_linear.weight"] = (3, self.W // 2)
            shp["rgb_linear.bias"] = (3,)
        else:
            shp["output_linear.weight"] = (self.output_ch, self.W)
            shp["output_linear.bias"] = (self.output_ch,)
        if self.semantic_num_classes:
            shp["semantic_linear.0.weight"] = (self.W // 2, self.W)
            shp["semantic_linear.0.bias"] = (self.W // 2,)
            shp["semantic_linear.1.weight"] = (self.semantic_num_classes, self.W // 2)
            shp["semantic_linear.1.bias"] = (self.semantic_num_classes,)
        return shp


def init_params(spec: MLPSpec, seed: int, dtype=torch.float32) -> Dict[str, Tensor]:
    """nn.Linear's default init (kaiming_uniform(a=sqrt 5) == U(-1/sqrt(fan_in), +)
    for both weight and bias), drawn from an explicit generator so that the same
    parameters can be rebuilt on the GPU box without the reference."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    shapes = spec.param_shapes()
    for name, shp in shapes.items():
        wname = name.rsplit(".", 1)[0] + ".weight"
        fan_in = shapes[wname][1]
        bound = 1.0 / math.sqrt(fan_in)
        out[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return out


def mlp_forward(p: Dict[str, Tensor], x: Tensor, spec: MLPSpec) -> Tensor:
    """run_nerf_helpers.py:113-145.  ``x`` is [..., input_ch + input_ch_views].

    Hidden stack with ReLU after every pts layer; after layer i in ``skips``
    the encoded position is concatenated IN FRONT of the hidden vector (:119-120).
    With view directions: alpha from the last hidden vector, `feature` (no
    activation) concatenated with the encoded direction, one W/2 ReLU layer,
    rgb; output is [rgb, alpha] (:122-140).  Without: one output layer (:143)."""
    x_pts, x_dir = torch.split(x, [spec.input_ch, spec.input_ch_views], dim=-1)
    h = x_pts
    for i in range(spec.D):
        h = torch.relu(h @ p[f"pts_linears.{i}.weight"].T + p[f"pts_linears.{i}.bias"])
        if i in spec.skips:
            h = torch.cat([x_pts, h], dim=-1)
    if not spec.use_viewdirs:
        return h @ p["output_linear.weight"].T + p["output_linear.bias"]
    sigma = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
    feat = h @ p["feature_linear.weight"].T + p["feature_linear.bias"]
    hv = torch.cat([feat, x_dir], dim=-1)
    hv = torch.relu(hv @ p["views_linears.0.weight"].T + p["views_linears.0.bias"])
    rgb = hv @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
    out = torch.cat([rgb, sigma], dim=-1)
    if spec.semantic_num_classes:
        sem = feat @ p["semantic_linear.0.weight"].T + p["semantic_linear.0.bias"]
        sem = sem @ p["semantic_linear.1.weight"].T + p["semantic_linear.1.bias"]
        out = torch.cat([out, sem], dim=-1)
    return out


def run_network(pts: Tensor, viewdirs: Optional[Tensor], p: Dict[str, Tensor], spec: MLPSpec,
                L_pts: int = 10, L_dir: int = 4, mlp_fn=None) -> Tensor:
    """run_nerf.py:60-74.  Encodes [N,S,3] points, broadcasts the per-ray view
    direction to every sample (:66-69) and applies the MLP.  (The reference's
    ``netchunk`` slicing (:50-57, :72) does not change results.)"""
    flat = pts.reshape(-1, pts.shape[-1])
    enc = posenc(flat, L_pts)
    if viewdirs is not None:
        d = viewdirs[:, None, :].expand(pts.shape).reshape(-1, 3)
        enc = torch.cat([enc, posenc(d, L_dir)], dim=-1)
    out = (mlp_fn or mlp_forward)(p, enc, spec)   # mlp_fn: tests plug in a bf16-emulating forward
    return out.reshape(*pts.shape[:-1], out.shape[-1])


# --------------------------------------------------------------------------- #
# R8  alpha compositing              run_nerf_helpers.py:542-595
# --------------------------------------------------------------------------- #
def raw2outputs(raw: Tensor, z_vals: Tensor, rays_d: Tensor, noise: Optional[Tensor] = None,
                white_bkgd: bool = False, semantic_loss: bool = False):
    """Returns (rgb_map, disp_map, acc_map, weights, depth_map[, semantic_class_preds]).

    ``noise`` is the ALREADY SCALED density noise (randn * raw_noise_std,
    run_nerf_helpers.py:563-565) or None.
    * interval lengths, last one 1e10, times |rays_d|            (:557-560)
    * colour = sigmoid(raw[..., :3])                              (:562)
    * alpha = 1 - exp(-relu(raw[..., 3] + noise) * dist)          (:555, :573)
    * weights = alpha * exclusive_cumprod(1 - alpha + 1e-10)      (:575)
    * rgb/depth/acc sums, disp = 1 / max(1e-10, depth / acc)      (:576-580)
    * optional white background                                   (:582-583)
    * semantic_loss: per-ray logits = UNWEIGHTED sum of raw[..., 4:] over the samples (:586-593)"""
    dt = raw.dtype
    delta = z_vals[..., 1:] - z_vals[..., :-1]
    delta = torch.cat([delta, torch.full_like(delta[..., :1], 1e10)], dim=-1)
    delta = delta * torch.linalg.norm(rays_d, dim=-1, keepdim=True)
    colour = torch.sigmoid(raw[..., :3])
    dens = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-torch.relu(dens) * delta)
    keep = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], dim=-1)
    trans = torch.cumprod(keep, dim=-1)[..., :-1]
    weights = alpha * trans
    rgb_map = (weights[..., None] * colour).sum(dim=-2)
    depth_map = (weights * z_vals).sum(dim=-1)
    acc_map = weights.sum(dim=-1)
    disp_map = 1.0 / torch.maximum(torch.full_like(depth_map, 1e-10), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    if semantic_loss:
        return rgb_map.to(dt), disp_map, acc_map, weights, depth_map, torch.sum(raw[..., 4:], dim=-2)
    return rgb_map.to(dt), disp_map, acc_map, weights, depth_map


# --------------------------------------------------------------------------- #
# R9  hierarchical sampling          run_nerf_helpers.py:497-540
# --------------------------------------------------------------------------- #
def pdf_to_cdf(weights: Tensor) -> Tensor:
    """run_nerf_helpers.py:499-502: +1e-5, normalise, cumsum, prepend 0."""
    w = weights + 1e-5
    pdf = w / w.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def invert_cdf(bins: Tensor, cdf: Tensor, u: Tensor):
    """run_nerf_helpers.py:523-538.  Returns (samples, inds).

    inds = searchsorted(cdf, u, right=True)  -> first index with cdf[i] > u
    below/above clamps, denom < 1e-5 -> 1, linear interpolation inside the bin."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = (inds - 1).clamp(min=0)
    hi = inds.clamp(max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    frac = (u - c_lo) / span
    return b_lo + frac * (b_hi - b_lo), inds


def sample_pdf(bins: Tensor, weights: Tensor, n_samples: int, det: bool = False,
               u: Optional[Tensor] = None) -> Tensor:
    """run_nerf_helpers.py:497-540.  ``u`` injects the uniform draws (the
    reference calls torch.rand at :509); det=True uses linspace(0,1,n) (:505-507)."""
    cdf = pdf_to_cdf(weights)
    if det:
        u = torch.linspace(0.0, 1.0, n_samples, dtype=cdf.dtype).expand(*cdf.shape[:-1], n_samples)
    elif u is None:
        u = torch.rand(*cdf.shape[:-1], n_samples, dtype=cdf.dtype)
    return invert_cdf(bins, cdf, u)[0]


def searchsorted_rows(a, v, side: str = "left"):
    """Batched row-wise search with row broadcast, the contract of the vendored
    extension (torchsearchsorted/src/torchsearchsorted/searchsorted.py:20-53,
    numpy restatement in .../utils.py:4-15).  numpy in, numpy int64 out."""
    import numpy as np
    a = np.asarray(a)
    v = np.asarray(v)
    rows = max(a.shape[0], v.shape[0])
    out = np.empty((rows, v.shape[1]), dtype=np.int64)
    for r in range(rows):
        out[r] = np.searchsorted(a[r if a.shape[0] > 1 else 0], v[r if v.shape[0] > 1 else 0], side=side)
    return out


# --------------------------------------------------------------------------- #
# R2  NDC warp                       run_nerf_helpers.py:320-337
# --------------------------------------------------------------------------- #
def ndc_rays(H: int, W: int, focal: float, near: float, rays_o: Tensor, rays_d: Tensor):
    """Moves origins to the near plane (:322-323) then applies the projective
    warp of the NeRF paper's appendix (:326-335)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    o = rays_o + t[..., None] * rays_d
    sx = -1.0 / (W / (2.0 * focal))
    sy = -1.0 / (H / (2.0 * focal))
    # operation order of :326-333 kept literally ((sx * o_x) / o_z, not sx * (o_x / o_z)): bit-exact with the reference
    o_ndc = torch.stack([sx * o[..., 0] / o[..., 2], sy * o[..., 1] / o[..., 2], 1.0 + 2.0 * near / o[..., 2]], dim=-1)
    d_ndc = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - o[..., 0] / o[..., 2]),
                         sy * (rays_d[..., 1] / rays_d[..., 2] - o[..., 1] / o[..., 2]),
                         -2.0 * near / o[..., 2]], dim=-1)
    return o_ndc, d_ndc


# --------------------------------------------------------------------------- #
# R4  per-ray renderer               run_nerf.py:520-675
# --------------------------------------------------------------------------- #
@dataclass
class RenderRNG:
    """The four draws one render_rays call makes, in the reference's order
    (run_nerf.py:585 -> helpers:565 -> helpers:509 -> helpers:565 again).
    Any entry left None means "that draw does not happen" (perturb==0 or
    raw_noise_std==0)."""
    t_rand: Optional[Tensor] = None       # U[0,1)  [N, N_samples]
    noise0: Optional[Tensor] = None       # N(0,1)  [N, N_samples]
    u: Optional[Tensor] = None            # U[0,1)  [N, N_importance]
    noise1: Optional[Tensor] = None       # N(0,1)  [N, N_samples + N_importance]


def stratified_z(near: Tensor, far: Tensor, n_samples: int, t_rand: Optional[Tensor],
                 lindisp: bool = False) -> Tensor:
    """run_nerf.py:571-593.  near/far are [N,1]."""
    t = torch.linspace(0.0, 1.0, n_samples, dtype=near.dtype)
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    else:
        z = near * (1.0 - t) + far * t
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], dim=-1)
        lo = torch.cat([z[..., :1], mid], dim=-1)
        z = lo + (hi - lo) * t_rand
    return z


def render_rays(ray_batch: Tensor, p_coarse, spec_coarse: MLPSpec, p_fine, spec_fine: MLPSpec,
                N_samples: int, N_importance: int, rng: RenderRNG, raw_noise_std: float = 0.0,
                white_bkgd: bool = False, lindisp: bool = False, L_pts: int = 10, L_dir: int = 4,
                retraw: bool = True, mlp_fn=None, semantic_loss: bool = False) -> Dict[str, Tensor]:
    """run_nerf.py:520-675 (the network_fn-is-not-None / no alpha_model /
    no sigma_loss branch, i.e. what every shipped config runs).  semantic_loss adds
    sem_preds / sem_preds0 (:601-603, :629-630, :643-644, :652-653, :662-663)."""
    o, d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    vdir = ray_batch[:, -3:] if ray_batch.shape[-1] > 9 else None          # :567
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z = stratified_z(near, far, N_samples, rng.t_rand, lindisp)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]                    # :595
    raw = run_network(pts, vdir, p_coarse, spec_coarse, L_pts, L_dir, mlp_fn)
    n0 = None if rng.noise0 is None else rng.noise0 * raw_noise_std
    rgb, disp, acc, w, depth, *sem = raw2outputs(raw, z, d, n0, white_bkgd, semantic_loss)
    out: Dict[str, Tensor] = {}
    if N_importance > 0:
        rgb0, disp0, acc0, depth0, sem0 = rgb, disp, acc, depth, sem
        z_mid = 0.5 * (z[..., 1:] + z[..., :-1])                           # :632
        z_new = sample_pdf(z_mid, w[..., 1:-1], N_importance, det=(rng.u is None), u=rng.u)
        z_new = z_new.detach()                                             # :634
        z, _ = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)           # :636
        pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
        raw = run_network(pts, vdir, p_fine if p_fine is not None else p_coarse,
                          spec_fine if p_fine is not None else spec_coarse, L_pts, L_dir, mlp_fn)
        n1 = None if rng.noise1 is None else rng.noise1 * raw_noise_std
        rgb, disp, acc, w, depth, *sem = raw2outputs(raw, z, d, n1, white_bkgd, semantic_loss)
        if semantic_loss:
            out["sem_preds0"] = sem0[0]
        out.update(rgb0=rgb0, disp0=disp0, acc0=acc0, depth_map0=depth0,
                   z_std=torch.std(z_new, dim=-1, unbiased=False))         # :659
    out.update(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth)
    if semantic_loss:
        out["sem_preds"] = sem[0]
    out["weights"] = w          # not returned by the reference; kept for kernel checks
    out["z_vals"] = z
    if retraw:
        out["raw"] = raw
    return out


# --------------------------------------------------------------------------- #
# R1  render                         run_nerf.py:112-194
# --------------------------------------------------------------------------- #
def pack_rays(H: int, W: int, focal: float, rays_o: Tensor, rays_d: Tensor, ndc: bool = True,
              near: float = 0.0, far: float = 1.0, use_viewdirs: bool = True) -> Tensor:
    """run_nerf.py:145-183: unit view directions from the PRE-warp directions,
    NDC warp with near plane 1, then [o, d, near, far, viewdirs]."""
    cols: List[Tensor] = []
    vd = None
    if use_viewdirs:
        vd = rays_d / torch.linalg.norm(rays_d, dim=-1, keepdim=True)
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, focal, 1.0, rays_o, rays_d)
    cols = [rays_o, rays_d, near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])]
    if vd is not None:
        cols.append(vd)
    return torch.cat(cols, dim=-1)


# --------------------------------------------------------------------------- #
# patch loss   loss.py:55-133
# --------------------------------------------------------------------------- #
# ---------------------------------------------------------------------------
# Ray generation
# ---------------------------------------------------------------------------
def pixel_rays(H: int, W: int, focal, c2w, x, y):
    """Rays through pixel positions (x, y) (numpy arrays of one floating dtype T): the shared arithmetic of
    get_rays_np (run_nerf_helpers.py:285-300), get_rays_by_coord_np (:303-318) and the crop generators (:430-494),
    every operation rounded once in T, the 3-term dot product summed left to right as numpy / CPU torch do."""
    T = x.dtype
    R = np.asarray(c2w)[:3, :4].astype(T)
    f = T.type(focal)
    d0 = (x - T.type(W * .5)) / f
    d1 = -((y - T.type(H * .5)) / f)
    d2 = -np.ones_like(d0)
    d = np.stack([(d0 * R[k, 0] + d1 * R[k, 1]) + d2 * R[k, 2] for k in range(3)], -1).astype(T)
    o = np.broadcast_to(R[:, 3], d.shape).astype(T)
    return o, d


def get_rays_np(H: int, W: int, focal, c2w):
    """run_nerf_helpers.py:285-300: all pixels of one camera, [H, W, 3] each, fp32."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    return pixel_rays(H, W, focal, c2w, i, j)


def get_rays_by_coord_np(H: int, W: int, focal, c2w, coords):
    """run_nerf_helpers.py:303-318: coords[N, 2] = (x, y), result in the coordinates' dtype."""
    c = np.asarray(coords)
    if c.dtype not in (np.float32, np.float64):
        c = c.astype(np.float64)
    return pixel_rays(H, W, focal, c2w, c[:, 0], c[:, 1])


def rays_cropped_feature_loss_new(H: int, W: int, focal, c2w, nH: int, nW: int, gradH: int, gradW: int, start_w: int,
                                  start_h: int, perm):
    """run_nerf_helpers.py:430-494 with the three random draws (crop corner :436-437, randperm :466) injected."""
    perm = np.asarray(perm).astype(np.int64)
    row, col = perm // nW, perm % nW
    o, d = pixel_rays(H, W, focal, c2w, (start_w + col).astype(np.float32), (start_h + row).astype(np.float32))
    pts = np.stack([row, col], -1)
    k = gradH * gradW
    return [o[:k], d[:k], pts[:k]], [o[k:], d[k:], pts[k:]], [start_w, start_w + nW - 1, start_h, start_h + nH - 1]


def inverse_depth_smoothness(idepth: Tensor, image: Tensor) -> Tensor:
    """InverseDepthSmoothnessLoss.forward (loss.py:87-133): forward differences a[.., j] - a[.., j+1] (:78-85),
    image weights exp(-mean_c |d image|) (:121-124), loss = mean|d_x idepth * w_x| + mean|d_y idepth * w_y| (:127-129).
    idepth (N,1,H,W), image (N,3,H,W)."""
    gx = lambda t: t[:, :, :, :-1] - t[:, :, :, 1:]        # noqa: E731
    gy = lambda t: t[:, :, :-1, :] - t[:, :, 1:, :]        # noqa: E731
    wx = torch.exp(-torch.mean(torch.abs(gx(image)), dim=1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(gy(image)), dim=1, keepdim=True))
    return torch.mean(torch.abs(gx(idepth) * wx)) + torch.mean(torch.abs(gy(idepth) * wy))


# --------------------------------------------------------------------------- #
# R11 loss assembly                  run_nerf.py:1451-1466, :1500-1536, :1759-1761
# --------------------------------------------------------------------------- #
def train_loss(out: Dict[str, Tensor], n_rgb: int, target_rgb: Tensor, target_depth: Optional[Tensor],
               depth_lambda: float = 0.0, depth_importance: float = 1.0,
               ray_weights: Optional[Tensor] = None, mode: str = "mse",
               target_semantic: Optional[Tensor] = None, semantic_lambda: float = 0.0) -> Dict[str, Tensor]:
    """RGB rays come first, depth rays after them (:1409-1411).  Colour losses use
    the RGB rays only (:1455-1462, :1500, :1759-1761); the depth loss supervises
    the FINE depth of the depth rays only (:1461, :1503-1524).

    mode: "mse" (:1524), "weighted" (:1517), "weighted_norm" (:1520),
    "relative" (:1522).  target_semantic (class index per RGB ray): cross-entropy of
    the fine and coarse per-ray logits of the RGB rays, times semantic_lambda (:1458-1460, :1541-1548)."""
    rgb = out["rgb_map"][:n_rgb]
    img = torch.mean((rgb - target_rgb) ** 2)
    res = {"img_loss": img}
    loss = img
    if target_depth is not None:
        dcol = out["depth_map"][n_rgb:]
        if mode == "weighted":
            dl = torch.mean(((dcol - target_depth) ** 2) * ray_weights)
        elif mode == "weighted_norm":
            dl = torch.mean((((dcol - target_depth) / torch.max(target_depth)) ** 2) * ray_weights)
        elif mode == "relative":
            dl = torch.mean(((dcol - target_depth) / (target_depth + 1e-16)) ** 2)
        else:
            dl = torch.mean((dcol - target_depth) ** 2)
        res["depth_loss"] = dl
        loss = loss + depth_lambda * depth_importance * dl
    if target_semantic is not None:
        sl = torch.nn.functional.cross_entropy(out["sem_preds"][:n_rgb], target_semantic)
        sl0 = torch.nn.functional.cross_entropy(out["sem_preds0"][:n_rgb], target_semantic) if "sem_preds0" in out else 0.0
        res["semantic_loss"], res["semantic_loss0"] = sl, sl0
        loss = loss + semantic_lambda * (sl + sl0)
    if "rgb0" in out:
        img0 = torch.mean((out["rgb0"][:n_rgb] - target_rgb) ** 2)
        res["img_loss0"] = img0
        loss = loss + img0
    res["loss"] = loss
    return res


# --------------------------------------------------------------------------- #
# Synthetic workloads; shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------- #
def synth_rays(n_rays: int, seed: int, H: int = 378, W: int = 504, focal: float = 407.6):
    """LLFF-shaped rays: forward-facing cameras near the origin looking down -z
    with small pose jitter; pixels uniformly random.  Mirrors get_rays_np
    (run_nerf_helpers.py:285-300) for the camera model.  Returns world-space
    (rays_o, rays_d) float32 [n,3]."""
    g = torch.Generator().manual_seed(seed)
    px = torch.rand(n_rays, generator=g) * (W - 1)
    py = torch.rand(n_rays, generator=g) * (H - 1)
    dirs = torch.stack([(px - W * 0.5) / focal, -(py - H * 0.5) / focal, -torch.ones_like(px)], dim=-1)
    ang = (torch.rand(n_rays, 3, generator=g) - 0.5) * (10.0 * math.pi / 180.0)
    cx, sx = torch.cos(ang[:, 0]), torch.sin(ang[:, 0])
    cy, sy = torch.cos(ang[:, 1]), torch.sin(ang[:, 1])
    # small rotations about x then y
    dx = dirs[:, 0]
    dy = dirs[:, 1] * cx - dirs[:, 2] * sx
    dz = dirs[:, 1] * sx + dirs[:, 2] * cx
    rays_d = torch.stack([dx * cy + dz * sy, dy, -dx * sy + dz * cy], dim=-1)
    t = torch.rand(n_rays, 3, generator=g)
    rays_o = torch.stack([(t[:, 0] - 0.5) * 0.6, (t[:, 1] - 0.5) * 0.6, (t[:, 2] - 0.5) * 0.1], dim=-1)
    return rays_o.float(), rays_d.float()


def synth_targets(n_rgb: int, n_depth: int, seed: int):
    """target_s ~ U(0,1)^3; NDC depth targets ~ U(0.25,1) with ~15 % 'sky' rays at
    1-1e-7 (load_llff.py:499,:521; Kitti360Dataset_new.py:191,:213)."""
    g = torch.Generator().manual_seed(seed + 1)
    tgt = torch.rand(n_rgb, 3, generator=g)
    dep = 0.25 + 0.75 * torch.rand(n_depth, generator=g)
    sky = torch.rand(n_depth, generator=g) < 0.15
    dep = torch.where(sky, torch.full_like(dep, 1.0 - 1e-7), dep)
    return tgt.float(), dep.float()


def synth_rng(n_rays: int, n_samples: int, n_importance: int, seed: int,
              perturb: bool = True, noise: bool = True) -> RenderRNG:
    g = torch.Generator().manual_seed(seed + 2)
    r = RenderRNG()
    if perturb:
        r.t_rand = torch.rand(n_rays, n_samples, generator=g)
    if noise:
        r.noise0 = torch.randn(n_rays, n_samples, generator=g)
    if perturb and n_importance > 0:
        r.u = torch.rand(n_rays, n_importance, generator=g)
    if noise and n_importance > 0:
        r.noise1 = torch.randn(n_rays, n_samples + n_importance, generator=g)
    return r


def trained_like(p: Dict[str, Tensor], sigma_bias: float = 2.0) -> Dict[str, Tensor]:
    """'Trained-like' variant: a positive density bias so that
    transmittance saturates and the coarse weights are peaky."""
    q = {k: v.clone() for k, v in p.items()}
    if "alpha_linear.bias" in q:
        q["alpha_linear.bias"] += sigma_bias
    return q
