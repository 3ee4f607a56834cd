"""Helpers shared by the `-m gpu` parity tests (test infrastructure, not product code)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import nerf_oracle as O  # noqa: E402

SLAB = 16384


def dn():
    import dlnerf_b200
    return dlnerf_b200


def report(name, got, ref, atol=0.0, rtol=0.0, quiet=False):
    """Print and return (max_abs_err, max_ref); asserts |got-ref| <= atol + rtol*|ref| on finite entries."""
    g = torch.as_tensor(got).detach().double().cpu()
    r = torch.as_tensor(ref).detach().double().cpu()
    assert g.shape == r.shape, (name, g.shape, r.shape)
    fin = torch.isfinite(r)
    assert torch.equal(torch.isfinite(g), fin), "%s: non-finite pattern differs" % name
    err = (g[fin] - r[fin]).abs()
    bound = atol + rtol * r[fin].abs()
    mx = err.max().item() if err.numel() else 0.0
    if not quiet:
        print("  %-34s max|err| %.3e  max|ref| %.3e  (atol %.1e rtol %.1e)" % (
            name, mx, r[fin].abs().max().item() if err.numel() else 0.0, atol, rtol))
    bad = err > bound
    assert not bad.any(), "%s: %d / %d entries out of tolerance, worst %.3e" % (name, int(bad.sum()), err.numel(), mx)
    return mx


def rel_l2(got, ref):
    g = torch.as_tensor(got).detach().double().cpu().flatten()
    r = torch.as_tensor(ref).detach().double().cpu().flatten()
    return ((g - r).norm() / (r.norm() + 1e-30)).item()


def cosine(got, ref):
    g = torch.as_tensor(got).detach().double().cpu().flatten()
    r = torch.as_tensor(ref).detach().double().cpu().flatten()
    return (g @ r / (g.norm() * r.norm() + 1e-30)).item()


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


_UNSWZ = None


def unswizzle_index():
    """int16-element index of (row, col) inside a SWIZZLE_128B slab image (see csrc/common.cuh slab_off)."""
    global _UNSWZ
    if _UNSWZ is None:
        r = torch.arange(128)[:, None]
        c = torch.arange(64)[None, :]
        off = (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 3) ^ r) & 7) << 4) + ((c & 7) << 1)
        _UNSWZ = (off // 2).long()
    return _UNSWZ


def read_stash(stash_u8, n_tiles, slots):
    """uint8 stash -> float tensor [n_tiles, slots, 128, 64]."""
    raw = stash_u8.detach().cpu().view(torch.int16).reshape(n_tiles, slots, SLAB // 2)
    idx = unswizzle_index().reshape(-1)
    vals = raw[:, :, idx].reshape(n_tiles, slots, 128, 64)
    return vals.contiguous().view(torch.bfloat16).float()


def stash_rows(st, slot, nslab, P):
    """[n_tiles, slots, 128, 64] -> [P, 64*nslab] (points x features) for `nslab` consecutive slots."""
    n_tiles = st.shape[0]
    x = st[:, slot:slot + nslab]                       # [T, nslab, 128, 64]
    x = x.permute(0, 2, 1, 3).reshape(n_tiles * 128, nslab * 64)
    return x[:P]


def mlp_reference(params, x, spec: O.MLPSpec, emulate_bf16=True):
    """Layer-by-layer torch reference of the MLP that mirrors the kernel's numerics: bf16 weights and
    bf16 activations between GEMM layers, fp32 accumulation, fp32 heads (alpha / rgb / output) on the
    un-rounded fp32 activations.  Returns dict with every intermediate."""
    q = (lambda t: bf16r(t)) if emulate_bf16 else (lambda t: t)
    xp, xd = torch.split(x, [spec.input_ch, spec.input_ch_views], dim=-1)
    out = {}
    xp_q, xd_q = q(xp), q(xd)
    h_q = xp_q
    h32 = None
    for i in range(spec.D):
        w = q(params["pts_linears.%d.weight" % i])
        h32 = torch.relu(h_q @ w.T + params["pts_linears.%d.bias" % i])
        out["H%d" % i] = q(h32)
        h_q = q(h32)
        if i in spec.skips:
            h_q = torch.cat([xp_q, h_q], -1)
    if not spec.use_viewdirs:
        out["raw"] = h32 @ params["output_linear.weight"].T + params["output_linear.bias"]
        return out
    sigma = h32 @ params["alpha_linear.weight"].T + params["alpha_linear.bias"]
    # feature_linear folded into views_linears (plan.build_plan, fold_feature): M = W_v1 W_f formed in fp32 and
    # rounded to bf16 once, b' = W_v1 b_f + b_v in fp32; the feature vector itself is never materialised
    Wv, W = params["views_linears.0.weight"], spec.W
    M = Wv[:, :W] @ params["feature_linear.weight"]
    b2 = Wv[:, :W] @ params["feature_linear.bias"] + params["views_linears.0.bias"]
    hv32 = torch.relu(h_q @ q(M).T + xd_q @ q(Wv[:, W:]).T + b2)
    out["HV"] = q(hv32)
    rgb = hv32 @ params["rgb_linear.weight"].T + params["rgb_linear.bias"]
    out["raw"] = torch.cat([rgb, sigma], -1)
    return out


def _ste(t):
    """bf16 rounding with a straight-through gradient."""
    return t + (bf16r(t) - t).detach()


def mlp_forward_emulated(params, x, spec: O.MLPSpec):
    """Differentiable twin of mlp_reference (same numerics as the kernels, straight-through rounding):
    autograd through it gives the gradient of the function the kernels actually evaluate, i.e. with the
    ReLU masks of the bf16 forward pass.  Drop-in for oracle.mlp_forward (mlp_fn=...)."""
    xp, xd = torch.split(x, [spec.input_ch, spec.input_ch_views], dim=-1)
    xp_q, xd_q = _ste(xp), _ste(xd)
    h_q, h32 = xp_q, None
    for i in range(spec.D):
        h32 = torch.relu(h_q @ _ste(params["pts_linears.%d.weight" % i]).T + params["pts_linears.%d.bias" % i])
        h_q = _ste(h32)
        if i in spec.skips:
            h_q = torch.cat([xp_q, h_q], -1)
    if not spec.use_viewdirs:
        return h32 @ params["output_linear.weight"].T + params["output_linear.bias"]
    sigma = h32 @ params["alpha_linear.weight"].T + params["alpha_linear.bias"]
    # feature_linear is folded into views_linears by the kernels' plan (plan.build_plan, fold_feature): the bf16
    # operand is M = W_v1 W_f (formed in fp32, rounded once), the bias b' = W_v1 b_f + b_v stays fp32
    Wv = params["views_linears.0.weight"]
    W = spec.W
    M = Wv[:, :W] @ params["feature_linear.weight"]
    b2 = Wv[:, :W] @ params["feature_linear.bias"] + params["views_linears.0.bias"]
    hv32 = torch.relu(h_q @ _ste(M).T + xd_q @ _ste(Wv[:, W:]).T + b2)
    rgb = hv32 @ params["rgb_linear.weight"].T + params["rgb_linear.bias"]
    out = torch.cat([rgb, sigma], -1)
    if spec.semantic_num_classes:
        # semantic head (csrc/semantic_kernels.cu): fp32 folded operands Sw = W_s2 W_s1 W_f, sc applied to the bf16
        # activations the chain keeps for wgrad
        A = params["semantic_linear.0.weight"] @ params["feature_linear.weight"]
        a = params["semantic_linear.0.weight"] @ params["feature_linear.bias"] + params["semantic_linear.0.bias"]
        Sw = params["semantic_linear.1.weight"] @ A
        sc = params["semantic_linear.1.weight"] @ a + params["semantic_linear.1.bias"]
        out = torch.cat([out, h_q @ Sw.T + sc], -1)
    return out


def compare_grads(named_got, ref_emul, ref_fp32, tag=""):
    """Print per-tensor cosine / rel-L2 against both references; return worst numbers and aggregates."""
    stats = dict(worst_cos_e=1.0, worst_l2_e=0.0, worst_cos_f=1.0, worst_l2_f=0.0)
    num_e = num_f = den_e = den_f = 0.0
    for name, g in named_got:
        if ref_fp32.get(name) is None:
            continue
        ce, le = cosine(g, ref_emul[name]), rel_l2(g, ref_emul[name])
        cf, lf = cosine(g, ref_fp32[name]), rel_l2(g, ref_fp32[name])
        print("  %s%-26s vs bf16-emulated: cos %.5f relL2 %.3e | vs fp32: cos %.5f relL2 %.3e | |ref| %.3e" % (
            tag, name, ce, le, cf, lf, ref_fp32[name].norm().item()))
        stats["worst_cos_e"], stats["worst_l2_e"] = min(stats["worst_cos_e"], ce), max(stats["worst_l2_e"], le)
        stats["worst_cos_f"], stats["worst_l2_f"] = min(stats["worst_cos_f"], cf), max(stats["worst_l2_f"], lf)
        gd = torch.as_tensor(g).detach().double().cpu()
        num_e += float((gd - ref_emul[name].double()).pow(2).sum()); den_e += float(ref_emul[name].double().pow(2).sum())
        num_f += float((gd - ref_fp32[name].double()).pow(2).sum()); den_f += float(ref_fp32[name].double().pow(2).sum())
    stats["agg_e"], stats["agg_f"] = (num_e / den_e) ** 0.5, (num_f / den_f) ** 0.5
    print("  %saggregate rel-L2: vs bf16-emulated %.3e, vs fp32 %.3e" % (tag, stats["agg_e"], stats["agg_f"]))
    return stats


def make_net(D, use_viewdirs=True, seed=0, device="cuda", sigma_bias=0.0, input_ch_views=27, semantic=0):
    """(package NeRF on device, oracle params dict on CPU, MLPSpec) with identical parameters."""
    spec = O.MLPSpec(D=D, use_viewdirs=use_viewdirs, input_ch_views=input_ch_views, semantic_num_classes=semantic)
    p = O.init_params(spec, seed=3407 + D + seed)
    if sigma_bias and use_viewdirs:
        p = O.trained_like(p, sigma_bias)
    net = dn().NeRF(D=D, W=256, input_ch=63, input_ch_views=input_ch_views, output_ch=5, skips=[4],
                    use_viewdirs=use_viewdirs, semantic_num_classes=semantic or None)
    net.load_state_dict(p)
    return net.to(device), p, spec
