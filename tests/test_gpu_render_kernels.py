"""GPU parity tests (through the C ABI) of the HBM-bound kernels: stratified sampling, positional
encoding, compositing forward/backward (+fused loss), hierarchical sampling, batched search.
Checker = oracle/nerf_oracle.py (pinned to the reference by tests/golden/)."""
import os

import numpy as np
import pytest
import torch

from gpu_util import O, dn, report

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = torch.from_numpy


def test_library_reports_build():
    assert b"sm_100a" in dn().lib().dln_build_info()


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dn().ops.posenc(torch.zeros(4, 3), 4)


# ---------------------------------------------------------------------------------------- sampling
@pytest.mark.parametrize("lindisp,perturb", [(False, True), (False, False), (True, True)])
def test_stratified_z(lindisp, perturb):
    g = torch.Generator().manual_seed(1)
    N, S = 333, 64
    rb = torch.zeros(N, 11)
    rb[:, 6] = 0.5 + torch.rand(N, generator=g)           # near
    rb[:, 7] = rb[:, 6] + 1.0 + torch.rand(N, generator=g)  # far
    tr = torch.rand(N, S, generator=g) if perturb else None
    ref = O.stratified_z(rb[:, 6:7], rb[:, 7:8], S, tr, lindisp)
    got = dn().ops.stratified_z(rb.to(DEV), S, None if tr is None else tr.to(DEV), lindisp)
    report("stratified_z", got, ref, atol=2e-7, rtol=2e-7)
    # NDC case near=0 far=1 is bit exact
    rb[:, 6], rb[:, 7] = 0.0, 1.0
    ref = O.stratified_z(rb[:, 6:7], rb[:, 7:8], S, tr, False)
    got = dn().ops.stratified_z(rb.to(DEV), S, None if tr is None else tr.to(DEV), False)
    assert torch.equal(got.cpu(), ref), "NDC stratified depths must match bit for bit"


def test_posenc_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "embed.npz"))
    x = T(g["x"]).to(DEV)
    report("posenc L=10 (golden)", dn().ops.posenc(x, 10), g["e10"], atol=2e-6)
    report("posenc L=4 (golden)", dn().ops.posenc(x, 4), g["e4"], atol=1e-6)
    fn, dim = dn().get_embedder(10, 0)
    assert dim == 63 and fn(x).shape == (x.shape[0], 63)
    ident, d3 = dn().get_embedder(10, -1)
    assert d3 == 3 and torch.equal(ident(x), x)
    e = dn().ops.posenc(torch.zeros(0, 3, device=DEV), 10)
    assert e.shape == (0, 63)


# ---------------------------------------------------------------------------------------- ray packing
@pytest.mark.parametrize("ndc,vd,near,far", [(True, True, 0., 1.), (False, True, 2., 6.), (True, False, 0., 1.)])
def test_pack_rays_matches_reference_packing(ndc, vd, near, far):
    """dln_pack_rays == render()'s packing (run_nerf.py:145-183 + ndc_rays) as restated by the pinned oracle:
    the kernel performs the reference's operations in the reference's order, each rounded separately, so the NDC
    origins / directions and the near / far columns are bit-exact; the unit view directions may differ in the last
    ulp (torch reduces the norm in a different association)."""
    Hh, Ww, foc = 378, 504, 407.6
    ro, rd = O.synth_rays(1000, seed=11)
    ref = O.pack_rays(Hh, Ww, foc, ro, rd, ndc=ndc, near=near, far=far, use_viewdirs=vd)
    got = dn().ops.pack_rays(Hh, Ww, foc, ro.to(DEV), rd.to(DEV), ndc, near, far, vd)
    assert got.shape == ref.shape
    assert torch.equal(got[:, :8].cpu(), ref[:, :8]), "o', d', near, far must be bit-exact"
    if vd:
        report("unit viewdirs", got[:, 8:], ref[:, 8:], atol=1.2e-7)
    # [2, N, 3] stacked input of the training loop and an empty batch
    rb = dn().pack_ray_batch(Hh, Ww, foc, torch.stack([ro, rd], 0).to(DEV), ndc, near, far, vd)
    assert torch.equal(rb, got)
    assert dn().ops.pack_rays(Hh, Ww, foc, ro[:0].to(DEV), rd[:0].to(DEV), ndc, near, far, vd).shape == (0, ref.shape[1])


# ---------------------------------------------------------------------------------------- compositing
def _composite_inputs(N, S, seed, C=4):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(N, S, C, generator=g) * 2.0
    z = torch.sort(torch.rand(N, S, generator=g), dim=-1)[0]
    rd = torch.randn(N, 3, generator=g) * 2.0
    nz = torch.randn(N, S, generator=g)
    return raw, z, rd, nz


@pytest.mark.parametrize("tag,std,wb", [("plain", 0.0, False), ("noise", 1.0, False), ("white", 0.5, True)])
def test_composite_forward_golden(golden_dir, tag, std, wb):
    g = np.load(os.path.join(golden_dir, "raw2outputs.npz"))
    raw, z, rd, nz = (T(g[k]).to(DEV) for k in ("raw", "z", "rays_d", "noise"))
    out = dn().ops.composite(raw, z, rd, nz if std > 0 else None, std, wb)
    for name, a in zip(["rgb", "disp", "acc", "weights", "depth"], out):
        report("composite/%s %s" % (tag, name), a, g["%s_%s" % (tag, name)], atol=2e-6, rtol=2e-5)


@pytest.mark.parametrize("N,S,C", [(1, 64, 4), (257, 64, 4), (100, 128, 4), (33, 40, 5), (17, 200, 4), (5, 2, 4)])
def test_composite_forward_backward_vs_oracle(N, S, C):
    raw, z, rd, nz = _composite_inputs(N, S, seed=N + S, C=C)
    g = torch.Generator().manual_seed(7)
    cot = [torch.randn(N, 3, generator=g), torch.randn(N, generator=g) * 1e-3, torch.randn(N, generator=g),
           torch.randn(N, S, generator=g), torch.randn(N, generator=g)]
    # fp64 oracle = ground truth
    raw64 = raw.double().requires_grad_(True)
    ref = O.raw2outputs(raw64, z.double(), rd.double(), nz.double() * 0.7, False)
    sum((r * c.double()).sum() for r, c in zip(ref, cot)).backward()
    rawg = raw.to(DEV).requires_grad_(True)
    out = dn().ops.composite(rawg, z.to(DEV), rd.to(DEV), nz.to(DEV), 0.7, False)
    for name, a, b in zip(["rgb", "disp", "acc", "weights", "depth"], out, ref):
        report("fwd %s [%d,%d,%d]" % (name, N, S, C), a, b, atol=3e-6, rtol=1e-4)
    sum((o * c.to(DEV)).sum() for o, c in zip(out, cot)).backward()
    gref = raw64.grad
    report("bwd d_raw [%d,%d,%d]" % (N, S, C), rawg.grad, gref, atol=1e-5 * gref.abs().max().item(), rtol=2e-3)


def test_composite_backward_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "raw2outputs.npz"))
    raw = T(g["raw"]).to(DEV).requires_grad_(True)
    out = dn().ops.composite(raw, T(g["z"]).to(DEV), T(g["rays_d"]).to(DEV), T(g["noise"]).to(DEV), 1.0, False)
    ((out[0] * T(g["c_rgb"]).to(DEV)).sum() + (out[4] * T(g["c_dep"]).to(DEV)).sum()
     + (out[2] * T(g["c_acc"]).to(DEV)).sum() + (out[3] * T(g["c_w"]).to(DEV)).sum()).backward()
    ref = g["draw"]
    report("d_raw (reference autograd, fp32)", raw.grad, ref, atol=2e-5 * np.abs(ref).max(), rtol=5e-3)


def test_composite_zero_rays_and_nan_disp():
    d = dn()
    out = d.ops.composite(torch.zeros(0, 64, 4, device=DEV), torch.zeros(0, 64, device=DEV),
                          torch.zeros(0, 3, device=DEV), None, 0.0, False)
    assert out[0].shape == (0, 3)
    # acc == 0 -> disp is NaN in the reference (0/0 inside max), SURVEY appendix A #12
    raw = torch.full((2, 8, 4), -5.0, device=DEV)
    z = torch.linspace(0, 1, 8, device=DEV).expand(2, 8).contiguous()
    rgb, disp, acc, w, depth = d.ops.composite(raw, z, torch.ones(2, 3, device=DEV), None, 0.0, False)
    assert (acc == 0).all() and torch.isnan(disp).all()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_fused_loss_backward(mode):
    """d raw from the fused loss kernel == autograd of the oracle's train_loss through raw2outputs."""
    N, S, n_rgb = 96, 64, 40
    raw, z, rd, nz = _composite_inputs(N, S, seed=5)
    tgt, dep = O.synth_targets(n_rgb, N - n_rgb, seed=3)
    rw = 0.5 + torch.rand(N - n_rgb)
    lam, imp = 0.01, 0.37
    names = {0: "mse", 1: "weighted", 2: "weighted_norm", 3: "relative"}
    raw64 = raw.double().requires_grad_(True)
    r = O.raw2outputs(raw64, z.double(), rd.double(), nz.double(), False)
    res = O.train_loss({"rgb_map": r[0], "depth_map": r[4]}, n_rgb, tgt.double(), dep.double(), lam, imp,
                       rw.double(), names[mode])
    res["loss"].backward()
    sums = torch.zeros(2, device=DEV)
    n_dep = N - n_rgb
    d_raw = dn().ops.composite_bwd_fused_loss(raw.to(DEV), z.to(DEV), rd.to(DEV), nz.to(DEV), 1.0, False,
                                              tgt.to(DEV), dep.to(DEV), rw.to(DEV), n_rgb, 2.0 / (3 * n_rgb),
                                              2.0 * lam * imp / n_dep, mode, float(dep.max()), sums)
    gref = raw64.grad
    report("fused-loss d_raw mode %d" % mode, d_raw, gref, atol=1e-5 * gref.abs().max().item(), rtol=2e-3)
    report("img_loss", sums[0] / (3 * n_rgb), res["img_loss"], rtol=1e-4)
    report("depth_loss", sums[1] / n_dep, res["depth_loss"], rtol=1e-4)


# ---------------------------------------------------------------------------------------- sample_pdf
def _check_samples(tag, got, ref, bin_width):
    """End-to-end samples against the reference path.  The inversion divides by the cdf step (clamped at
    1e-5, helpers:536), so a 1-ulp cdf difference is amplified by up to 1e5 * bin width and the
    `denom < 1e-5 -> 1` rule is discontinuous; the reference's own CPU and CUDA builds differ the same way.
    Hence: 99 % of the samples within 2e-5, all of them within one bin."""
    err = (got.detach().cpu() - torch.as_tensor(ref)).abs()
    frac = (err <= 2e-5).float().mean().item()
    print("  %-34s within 2e-5: %.4f  max|err| %.3e" % (tag, frac, err.max().item()))
    assert frac >= 0.99 and err.max().item() <= bin_width, tag


def test_sample_pdf_golden_and_bit_exact_indices(golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_pdf.npz"))
    bins, w, u = T(g["bins"]).to(DEV), T(g["w"]).to(DEV), T(g["u"]).to(DEV)
    s, cdf, inds = dn().ops.sample_pdf(bins, w, 64, u, return_debug=True)
    report("cdf", cdf, g["cdf"], atol=5e-7)
    # bit-exact index search on the kernel's own cdf (torch.searchsorted == the reference's call, helpers:524)
    want = torch.searchsorted(cdf.cpu(), u.cpu().contiguous(), right=True)
    assert torch.equal(inds.cpu(), want), "searchsorted(right=True) indices differ"
    # and equal to the reference's indices wherever the two cdfs agree bit for bit
    same_rows = (cdf.cpu() == T(g["cdf"])).all(dim=1)
    print("  rows with bit-identical cdf: %d / %d" % (int(same_rows.sum()), cdf.shape[0]))
    assert torch.equal(inds.cpu()[same_rows], T(g["inds"])[same_rows])
    # given the same cdf, the inversion arithmetic (helpers:525-538) is bit exact
    s_from_cdf, _ = O.invert_cdf(T(g["bins"]), cdf.cpu(), T(g["u"]))
    assert torch.equal(s.cpu(), s_from_cdf), "interpolation differs from the reference formula on identical cdf"
    _check_samples("samples (rand u) vs reference", s, g["s_rand"], 1.0 / 16)
    s_det, cdf_d, inds_d = dn().ops.sample_pdf(bins, w, 64, None, return_debug=True)
    u_det = torch.linspace(0., 1., 64).expand(bins.shape[0], 64).contiguous()
    assert torch.equal(inds_d.cpu(), torch.searchsorted(cdf_d.cpu(), u_det, right=True)), "det u must be linspace"
    assert torch.equal(s_det.cpu(), O.invert_cdf(T(g["bins"]), cdf_d.cpu(), u_det)[0])
    _check_samples("samples (det) vs reference", s_det, g["s_det"], 1.0 / 16)
    # the drop-in signature
    s2 = dn().sample_pdf(bins, w, 64, det=True)
    assert torch.equal(s2, s_det)


@pytest.mark.parametrize("N,S,Ni", [(1, 64, 64), (300, 64, 64), (64, 64, 128), (50, 33, 17), (9, 128, 64)])
def test_importance_resample(N, S, Ni):
    g = torch.Generator().manual_seed(N * 7 + S)
    z = torch.sort(torch.rand(N, S, generator=g), dim=-1)[0]
    w = torch.rand(N, S, generator=g) ** 6
    w[0] = 0.0
    u = torch.rand(N, Ni, generator=g)
    zs, zm, cdf, inds = dn().ops.importance_resample(z.to(DEV), w.to(DEV), Ni, u.to(DEV), return_debug=True)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    report("cdf", cdf, O.pdf_to_cdf(w[:, 1:-1]), atol=5e-7)
    assert torch.equal(inds.cpu(), torch.searchsorted(cdf.cpu(), u, right=True))
    assert torch.equal(zs.cpu(), O.invert_cdf(mids, cdf.cpu(), u)[0]), "inversion must be bit exact on identical cdf"
    _check_samples("z_samples vs reference path", zs, O.sample_pdf(mids, w[:, 1:-1], Ni, u=u), (z[:, -1] - z[:, 0]).max().item())
    ref_m = torch.sort(torch.cat([z, zs.cpu()], -1), -1)[0]
    assert torch.equal(zm.cpu(), ref_m), "merged depths must equal sort(cat(z_vals, z_samples)) exactly"
    zs_det, zm_det, cdf_d, inds_d = dn().ops.importance_resample(z.to(DEV), w.to(DEV), Ni, None, return_debug=True)
    u_det = torch.linspace(0., 1., Ni).expand(N, Ni).contiguous()
    assert torch.equal(zs_det.cpu(), O.invert_cdf(mids, cdf_d.cpu(), u_det)[0])
    assert (zm_det[:, 1:] >= zm_det[:, :-1]).all()


@pytest.mark.parametrize("N", [1, 97, 4096 + 17])
def test_importance_resample_thread_per_ray_kernel(N, monkeypatch):
    """The 64 + 64 shape has two kernels (render_kernels.cu: warp per ray below 32 768 rays, thread per ray above):
    same bits from both, for injected u, the deterministic u and the in-kernel Philox draws, with the rays of a
    partial last warp included; and the thread-per-ray kernel alone against the reference arithmetic."""
    S = Ni = 64
    g = torch.Generator().manual_seed(1000 + N)
    z = torch.sort(torch.rand(N, S, generator=g), dim=-1)[0]
    w = torch.rand(N, S, generator=g) ** 6
    w[0] = 0.0
    w[N // 2, 5:] = 0.0                                            # a cdf that saturates early: denom < 1e-5 -> 1
    u = torch.rand(N, Ni, generator=g)
    u[0, :4] = torch.tensor([0.0, 1.0 - 2.0 ** -24, 0.5, 1e-38])   # ends of the range, a numerator div.rn's fast path rejects
    zd, wd, ud = z.to(DEV), w.to(DEV), u.to(DEV)
    st = dn().ops.RngState(DEV, 1234)
    out = {}
    for mode in ("warp", "thread"):
        monkeypatch.setenv("DLN_RESAMPLE", mode)
        out[mode] = (dn().ops.importance_resample(zd, wd, Ni, ud, return_debug=True),
                     dn().ops.importance_resample(zd, wd, Ni, None, return_debug=True),
                     None,
                     dn().ops.importance_resample(zd, wd, Ni, rng=(st, 7)))
    for case in (0, 1, 3):
        for a, b in zip(out["warp"][case], out["thread"][case]):
            assert torch.equal(a, b), "warp-per-ray and thread-per-ray kernels differ (case %d)" % case
    zs, zm, cdf, inds = out["thread"][0]
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    report("cdf (thread per ray)", cdf, O.pdf_to_cdf(w[:, 1:-1]), atol=5e-7)
    assert torch.equal(inds.cpu(), torch.searchsorted(cdf.cpu(), u, right=True))
    assert torch.equal(zs.cpu(), O.invert_cdf(mids, cdf.cpu(), u)[0]), "inversion must be bit exact on identical cdf"
    assert torch.equal(zm.cpu(), torch.sort(torch.cat([z, zs.cpu()], -1), -1)[0])
    # a weights view that is not 4 bytes past a 16-byte boundary takes the scalar-load instantiation
    w_off = torch.zeros(N * S + 3, device=DEV)[3:].view(N, S)
    w_off.copy_(wd)
    zs2, zm2 = dn().ops.importance_resample(zd, w_off, Ni, ud)
    assert torch.equal(zs2, zs) and torch.equal(zm2, zm)


def test_importance_resample_default_dispatch_at_config_e_size(monkeypatch):
    """32 768 + 50 rays (a config-E chunk): the default dispatch takes the thread-per-ray kernel there; its output equals
    the warp-per-ray kernel's and sort(cat(z_vals, z_samples)), with live Philox draws and with injected u."""
    N, S, Ni = 32768 + 50, 64, 64
    g = torch.Generator().manual_seed(77)
    z = torch.sort(torch.rand(N, S, generator=g), dim=-1)[0].to(DEV)
    w = (torch.rand(N, S, generator=g) ** 8).to(DEV)
    u = torch.rand(N, Ni, generator=g).to(DEV)
    st = dn().ops.RngState(DEV, 4321)
    monkeypatch.delenv("DLN_RESAMPLE", raising=False)
    zs_d, zm_d = dn().ops.importance_resample(z, w, Ni, u)
    zs_r, zm_r = dn().ops.importance_resample(z, w, Ni, rng=(st, 2))
    monkeypatch.setenv("DLN_RESAMPLE", "warp")
    zs_w, zm_w = dn().ops.importance_resample(z, w, Ni, u)
    zs_rw, zm_rw = dn().ops.importance_resample(z, w, Ni, rng=(st, 2))
    assert torch.equal(zs_d, zs_w) and torch.equal(zm_d, zm_w) and torch.equal(zs_r, zs_rw) and torch.equal(zm_r, zm_rw)
    assert torch.equal(zm_d, torch.sort(torch.cat([z, zs_d], -1), -1)[0])
    assert torch.equal(zm_r, torch.sort(torch.cat([z, zs_r], -1), -1)[0])
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    assert (zs_r >= mids[:, :1] - 1e-6).all() and (zs_r <= mids[:, -1:] + 1e-6).all()


@pytest.mark.parametrize("N,Ni,C", [(1, 64, 4), (301, 64, 4), (77, 40, 4), (130, 64, 7)])
def test_composite_resample_fused_launch(N, Ni, C):
    """dln_composite_resample_fwd (coarse raw2outputs + hierarchical resampling in one launch, what train_step calls)
    returns the bits of the two separate calls -- injected draws, in-kernel draws, and neither."""
    S = 64
    g = torch.Generator().manual_seed(5 * N + Ni)
    raw = (torch.randn(N, S, C, generator=g) * 2).to(DEV)
    z = torch.sort(torch.rand(N, S, generator=g), -1)[0].to(DEV)
    rd = torch.randn(N, 3, generator=g).to(DEV)
    noise = torch.randn(N, S, generator=g).to(DEV)
    u = torch.rand(N, Ni, generator=g).to(DEV)
    st = dn().ops.RngState(DEV, 99)
    for nz, uu, rn, ru, std in ((noise, u, None, None, 1.0), (None, None, (st, 3), (st, 5), 1.0), (None, None, None, None, 0.0)):
        sep = dn().ops.composite(raw, z, rd, nz, std, False, rng=rn)
        zs, zm = dn().ops.importance_resample(z, sep[3], Ni, uu, rng=ru)
        fused = dn().ops.composite_resample(raw, z, rd, nz, std, False, Ni, uu, rng_noise=rn, rng_u=ru)
        for name, a, b in zip(("rgb", "disp", "acc", "weights", "depth", "z_samples", "z_merged"), fused, list(sep) + [zs, zm]):
            assert torch.equal(a, b) or (torch.isnan(a) == torch.isnan(b)).all() and torch.equal(a.nan_to_num(), b.nan_to_num()), name
    with pytest.raises(ValueError):
        dn().ops.composite_resample(raw[:, :32], z[:, :32], rd, None, 0.0, False, Ni)


# ---------------------------------------------------------------------------------------- searchsorted
def test_searchsorted_grid_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "searchsorted.npz"))
    n = int(g["n"][0])
    for i in range(n):
        side = "left" if g["side%d" % i][0] == 0 else "right"
        out = dn().ops.searchsorted(T(g["a%d" % i]).to(DEV), T(g["v%d" % i]).to(DEV), side)
        assert out.dtype == torch.int64
        np.testing.assert_array_equal(out.cpu().numpy(), g["r%d" % i])
    print("  %d searchsorted cases bit-exact" % n)


def test_searchsorted_argument_errors():
    a = torch.zeros(3, 5, device=DEV)
    with pytest.raises(ValueError):
        dn().ops.searchsorted(a, torch.zeros(2, 5, device=DEV))
    with pytest.raises(ValueError):
        dn().ops.searchsorted(a[0], a)
    with pytest.raises(ValueError):
        dn().ops.searchsorted(a, a, side="middle")
    ka = dn().ops.searchsorted(torch.tensor([[0, .25, .25, .5, 1.]], device=DEV),
                               torch.tensor([[0, .25, .3, .5, 1.]], device=DEV), "right")
    assert ka.cpu().tolist() == [[1, 3, 3, 4, 5]]


def test_full_size_properties():
    """BASELINE config sizes (4096 rays, 64+64): size-independent properties."""
    N, S, Ni = 4096, 64, 64
    g = torch.Generator().manual_seed(11)
    raw = (torch.randn(N, S, 4, generator=g) * 2).to(DEV)
    z = torch.sort(torch.rand(N, S, generator=g), -1)[0].to(DEV)
    rd = torch.randn(N, 3, generator=g).to(DEV)
    rgb, disp, acc, w, depth = dn().ops.composite(raw, z, rd, None, 0.0, False)
    assert (w >= 0).all() and (acc <= 1 + 1e-5).all()
    report("acc == sum(weights)", acc, w.sum(-1), atol=1e-5)
    assert ((depth >= z[:, 0] * acc - 1e-5) & (depth <= z[:, -1] * acc + 1e-5)).all()
    assert (rgb >= -1e-6).all() and (rgb <= acc[:, None] + 1e-5).all()
    zs, zm = dn().ops.importance_resample(z, w, Ni, torch.rand(N, Ni, generator=g).to(DEV))
    assert (zm[:, 1:] >= zm[:, :-1]).all()
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    assert (zs >= mids[:, :1] - 1e-6).all() and (zs <= mids[:, -1:] + 1e-6).all()
    # idempotence: merging is a permutation of the inputs
    assert torch.equal(torch.sort(torch.cat([z, zs], -1), -1)[0], zm)


# ---------------------------------------------------------------------------------------- patch loss
def test_inverse_depth_smoothness_loss_golden_and_patch(golden_dir):
    """dn.InverseDepthSmoothnessLoss (dln_inv_depth_smooth_fwd/bwd) against the reference's loss.py:55-133: the
    golden fixture generated by the unmodified reference, then a KITTI-360-sized patch pair [2, ., 94, 352] against
    the pinned oracle in float64.  fp32 sums over 66 k terms: rtol 2e-6 on the value, 1e-6 x max|g| on gradients."""
    d = dn()
    crit = d.InverseDepthSmoothnessLoss()
    g = np.load(os.path.join(golden_dir, "inv_depth_smooth.npz"))
    idp = T(g["idepth"]).to(DEV).requires_grad_(True)
    img = T(g["image"]).to(DEV).requires_grad_(True)
    loss = crit(idp, img)
    (3.0 * loss).backward()
    report("loss (golden)", loss, g["loss"], rtol=2e-6)
    report("d loss / d idepth (golden)", idp.grad, 3.0 * g["g_idepth"], atol=1e-6 * np.abs(g["g_idepth"]).max() * 3)
    report("d loss / d image (golden)", img.grad, 3.0 * g["g_image"], atol=1e-6 * np.abs(g["g_image"]).max() * 3)
    gen = torch.Generator().manual_seed(3)
    idp = torch.rand(2, 1, 94, 352, generator=gen) + 0.01
    img = torch.rand(2, 3, 94, 352, generator=gen)
    i64, m64 = idp.double().requires_grad_(True), img.double().requires_grad_(True)
    ref = O.inverse_depth_smoothness(i64, m64)
    ref.backward()
    a, b = idp.to(DEV).requires_grad_(True), img.to(DEV).requires_grad_(True)
    out = crit(a, b)
    out.backward()
    report("loss (94x352 patch)", out, ref, rtol=2e-6)
    report("d idepth (94x352 patch)", a.grad, i64.grad, atol=1e-6 * i64.grad.abs().max().item())
    report("d image (94x352 patch)", b.grad, m64.grad, atol=1e-6 * m64.grad.abs().max().item())
    # only one input needs a gradient (the no-grad part of a patch); shape / type errors as in the reference
    a2 = idp.to(DEV).requires_grad_(True)
    crit(a2, img.to(DEV)).backward()
    report("d idepth, image without grad", a2.grad, i64.grad, atol=1e-6 * i64.grad.abs().max().item())
    with pytest.raises(ValueError):
        crit(idp.to(DEV)[0], img.to(DEV))
    with pytest.raises(TypeError):
        crit(idp.numpy(), img.to(DEV))
