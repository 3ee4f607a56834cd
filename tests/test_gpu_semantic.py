"""GPU parity of the semantic head (run_nerf_helpers.py:107-111, :126-127, :586-593; run_nerf.py:1541-1548;
csrc/semantic_kernels.cu) against the oracle, which tests/golden/semantic.npz pins bit-exactly to the reference.

Stated tolerances (bf16 trunk, fp32 head on the bf16 activations the chain keeps):
    per-point logits            abs <= 2e-2 + 1e-2 |ref|
    per-ray logits (sum of 64 / 128 samples)      abs <= 5e-2 + 1e-2 |ref|
    losses                      rel <= 2e-2
    gradients                   as in test_gpu_render_e2e.py: per tensor cosine >= 0.995 and aggregate rel-L2 <= 5e-3 (measured 1.4e-3)
                                against autograd through the bf16-emulating twin, aggregate rel-L2 <= 3e-2 against
                                the fp32 oracle
    fp32-only kernels (sample sums, cross-entropy)   rtol 1e-5 / atol 1e-6
"""
import os

import numpy as np
import pytest
import torch

from gpu_util import O, compare_grads, dn, make_net, mlp_forward_emulated, rel_l2, report

pytestmark = pytest.mark.gpu
DEV = "cuda"
K = 19
H, W, FOCAL = 94, 352, 138.14          # KITTI-360 shape (fern_dsnerf.txt:22, :50-51, :55)


def test_module_forward_and_backward_with_semantic_logits():
    """NeRF.forward(x) -> [..., 4 + K] like the reference module, all 28 gradients."""
    net, p, spec = make_net(8, seed=3, semantic=K)
    g = torch.Generator().manual_seed(5)
    x = torch.cat([O.posenc(torch.rand(300, 3, generator=g) * 2 - 1, 10),
                   O.posenc(torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1), 4)], -1)
    cot = torch.randn(300, 4 + K, generator=g)
    y = net(x.to(DEV))
    assert y.shape == (300, 4 + K)
    p32 = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y32 = O.mlp_forward(p32, x, spec)
    report("rgb / sigma", y[:, :4], y32[:, :4], atol=2e-2, rtol=1e-2)
    report("semantic logits", y[:, 4:], y32[:, 4:], atol=2e-2, rtol=1e-2)
    pe = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ye = mlp_forward_emulated(pe, x, spec)
    report("semantic logits vs bf16 twin", y[:, 4:], ye[:, 4:], atol=5e-3 * ye[:, 4:].abs().max().item() + 1e-3)
    (y * cot.to(DEV)).sum().backward()
    (y32 * cot).sum().backward()
    (ye * cot).sum().backward()
    st = compare_grads([(n, q.grad) for n, q in net.named_parameters()], {k: v.grad for k, v in pe.items()},
                       {k: v.grad for k, v in p32.items()})
    assert len([1 for _, q in net.named_parameters() if q.grad is not None]) == 28
    assert st["worst_cos_e"] >= 0.995 and st["agg_e"] <= 5e-3 and st["agg_f"] <= 3e-2


def test_raw2outputs_semantic_against_reference_golden(golden_dir):
    """Generic route: a caller-supplied raw[N, S, 4+K]; outputs of the UNMODIFIED reference."""
    g = np.load(os.path.join(golden_dir, "semantic.npz"))
    d = dn()
    raw = torch.from_numpy(g["r2o_raw"]).to(DEV).requires_grad_(True)
    out = d.raw2outputs(raw, torch.from_numpy(g["r2o_z"]).to(DEV), torch.from_numpy(g["r2o_rays_d"]).to(DEV),
                        semantic_loss=True)
    assert len(out) == 6
    report("semantic_class_preds (golden)", out[5], g["r2o_sem"], atol=1e-5, rtol=1e-5)
    report("rgb_map (golden, 4+K channels)", out[0], g["r2o_rgb"], atol=1e-5)
    report("depth_map (golden, 4+K channels)", out[4], g["r2o_depth"], atol=1e-5)
    cot = torch.randn(out[5].shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    ((out[5] * cot).sum() + out[0].sum()).backward()
    want = cot[:, None, :].expand(-1, raw.shape[1], -1)
    report("d raw[..., 4:] (broadcast)", raw.grad[..., 4:], want, atol=0.0)
    assert raw.grad[..., :4].abs().sum() > 0


def test_cross_entropy_kernel_against_torch():
    d = dn()
    g = torch.Generator().manual_seed(2)
    N, n_rgb = 301, 200
    x = torch.randn(N, K, generator=g) * 6
    t = torch.randint(0, K, (n_rgb,), generator=g)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(xr[:n_rgb], t)
    (0.3 * ref).backward()
    sums = torch.zeros(1, device=DEV)
    dsem = d.ops.semantic_ce(x.to(DEV), t.to(DEV), n_rgb, 0.3 / n_rgb, sums)
    report("cross-entropy", sums[0] / n_rgb, ref, rtol=1e-5, atol=1e-6)
    report("d logits", dsem, xr.grad, rtol=1e-4, atol=1e-7)
    assert torch.equal(dsem[n_rgb:], torch.zeros_like(dsem[n_rgb:]))


def _case(n_rgb, n_dep, seed, perturb=True, noise=True):
    net_c, pc, spec_c = make_net(4, seed=seed, sigma_bias=1.0, semantic=K)
    net_f, pf, spec_f = make_net(8, seed=seed + 1, sigma_bias=1.0, semantic=K)
    ro, rd = O.synth_rays(n_rgb + n_dep, seed=seed, H=H, W=W, focal=FOCAL)
    rng = O.synth_rng(n_rgb + n_dep, 64, 64, seed=seed, perturb=perturb, noise=noise)
    tgt, dep = O.synth_targets(n_rgb, n_dep, seed=seed)
    tsem = torch.randint(0, K, (n_rgb,), generator=torch.Generator().manual_seed(seed))
    return net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem


def _oracle(pc, spec_c, pf, spec_f, ro, rd, rng, tgt, dep, tsem, n_rgb, std, lam, imp, slam, mlp_fn=None):
    rb = O.pack_rays(H, W, FOCAL, ro, rd)
    pcg = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    pfg = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    out = O.render_rays(rb, pcg, spec_c, pfg, spec_f, 64, 64, rng, raw_noise_std=std, mlp_fn=mlp_fn, semantic_loss=True)
    res = O.train_loss(out, n_rgb, tgt, dep, depth_lambda=lam, depth_importance=imp, target_semantic=tsem,
                       semantic_lambda=slam)
    res["loss"].backward()
    return out, res, pcg, pfg


def _render(net_c, net_f, ro, rd, rng, std, perturb, query=None, retraw=True, chunk=1 << 20):
    d = dn()
    e_p, e_d = d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0]
    q = query or d.FusedQuery(e_p, e_d, 1 << 16, 10, 4, 0)
    inj = {k: getattr(rng, k).to(DEV) for k in ("t_rand", "noise0", "u", "noise1") if getattr(rng, k) is not None}
    kw = dict(network_query_fn=q, perturb=1.0 if perturb else 0.0, N_importance=64, network_fine=net_f, N_samples=64,
              network_fn=net_c, use_viewdirs=True, white_bkgd=False, raw_noise_std=std, ndc=True, semantic_loss=True,
              _rng=inj)
    return d.render(H, W, FOCAL, chunk=chunk, rays=torch.stack([ro, rd], 0).to(DEV), retraw=retraw, near=0., far=1., **kw)


def test_render_with_semantic_loss_and_backward_parity():
    """The reference's training iteration with semantic_loss = True (fern_dsnerf.txt:55-56): render, RGB + depth +
    cross-entropy (fine and coarse) loss, backward -- drop-in route against the oracle.  An odd ray count leaves the
    coarse pass with a half-filled last tile (padded rows must not receive the per-ray semantic gradient)."""
    n_rgb, n_dep = 161, 96
    lam, imp, slam = 0.01, 0.5, 0.01
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem = _case(n_rgb, n_dep, 61)
    ref, res, pcg, pfg = _oracle(pc, spec_c, pf, spec_f, ro, rd, rng, tgt, dep, tsem, n_rgb, 1.0, lam, imp, slam)
    rgb, disp, acc, depth, extras = _render(net_c, net_f, ro, rd, rng, 1.0, True)
    assert set(extras) == {"raw", "rgb0", "disp0", "acc0", "depth_map0", "z_std", "sem_preds", "sem_preds0"}
    assert extras["raw"].shape == (n_rgb + n_dep, 128, 4 + K) and extras["sem_preds"].shape == (n_rgb + n_dep, K)
    report("rgb_map", rgb, ref["rgb_map"], atol=2e-2)
    report("depth_map", depth, ref["depth_map"], atol=2e-2)
    report("sem_preds", extras["sem_preds"], ref["sem_preds"], atol=5e-2, rtol=1e-2)
    report("sem_preds0", extras["sem_preds0"], ref["sem_preds0"], atol=5e-2, rtol=1e-2)
    report("raw[..., 4:] (per-sample logits)", extras["raw"][..., 4:], ref["raw"][..., 4:], atol=2e-2, rtol=1e-2)
    report("sum of raw[..., 4:] == sem_preds", extras["raw"][..., 4:].sum(1), extras["sem_preds"], atol=2e-3, rtol=1e-4)
    d = dn()
    F = torch.nn.functional
    ts = tsem.to(DEV)
    loss = d.img2mse(rgb[:n_rgb], tgt.to(DEV)) + lam * imp * d.img2mse(depth[n_rgb:], dep.to(DEV)) \
        + slam * (F.cross_entropy(extras["sem_preds"][:n_rgb], ts) + F.cross_entropy(extras["sem_preds0"][:n_rgb], ts)) \
        + d.img2mse(extras["rgb0"][:n_rgb], tgt.to(DEV))
    report("loss", loss, res["loss"], rtol=2e-2)
    loss.backward()
    _, _, pce, pfe = _oracle(pc, spec_c, pf, spec_f, ro, rd, rng, tgt, dep, tsem, n_rgb, 1.0, lam, imp, slam,
                             mlp_forward_emulated)
    for net, p32, pem, tag in ((net_f, pfg, pfe, "fine   "), (net_c, pcg, pce, "coarse ")):
        st = compare_grads([(n, q.grad) for n, q in net.named_parameters()],
                           {k: v.grad for k, v in pem.items()}, {k: v.grad for k, v in p32.items()}, tag)
        assert st["worst_cos_e"] >= 0.995 and st["agg_e"] <= 5e-3, tag
        assert st["agg_f"] <= 3e-2, tag
        for n, q in net.named_parameters():
            if n.startswith("semantic_linear"):
                assert q.grad is not None and float(q.grad.abs().sum()) > 0, n


def test_semantic_only_loss_reaches_the_trunk():
    """A loss made of the cross-entropy alone: every gradient then flows through the per-ray row the dgrad chain
    adds to dH of the last trunk layer (DlnChainArgs.sem_g) and through the fold / unfold kernels."""
    n = 96
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem = _case(n, 0, 67, False, False)
    rgb, disp, acc, depth, extras = _render(net_c, net_f, ro, rd, rng, 0.0, False, retraw=False)
    assert "raw" not in extras
    loss = torch.nn.functional.cross_entropy(extras["sem_preds"], tsem.to(DEV))
    loss.backward()
    rb = O.pack_rays(H, W, FOCAL, ro, rd)
    grads = {}
    for tag, fn in (("emul", mlp_forward_emulated), ("fp32", None)):
        pfg = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
        out = O.render_rays(rb, pc, spec_c, pfg, spec_f, 64, 64, rng, mlp_fn=fn, semantic_loss=True)
        torch.nn.functional.cross_entropy(out["sem_preds"], tsem).backward()
        grads[tag] = {k: v.grad for k, v in pfg.items()}
    used = [(k, q.grad) for k, q in net_f.named_parameters()
            if not k.startswith(("views_linears", "rgb_linear", "alpha_linear"))]
    st = compare_grads(used, grads["emul"], grads["fp32"], "fine   ")
    assert st["worst_cos_e"] >= 0.995 and st["agg_e"] <= 5e-3 and st["agg_f"] <= 3e-2
    for k, q in net_f.named_parameters():            # the colour / density heads see no gradient from this loss
        if k.startswith(("views_linears", "rgb_linear", "alpha_linear")):
            assert float(q.grad.abs().max()) <= 1e-12, k


def test_generic_query_route_matches_fused_route_with_semantics():
    """A foreign network_query_fn (the reference's lambda): raw[N, S, 4+K] from NeRF.forward, summed by the
    sample-sum kernel -- same sem_preds as the fused route, ray chunks included."""
    n = 48
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem = _case(n, 0, 71, False, False)
    d = dn()
    e_p, e_d = d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0]
    generic = lambda inputs, viewdirs, fn: d.run_network(inputs, viewdirs, fn, embed_fn=e_p, embeddirs_fn=e_d,  # noqa: E731
                                                         netchunk=4096)
    with torch.no_grad():
        a = _render(net_c, net_f, ro, rd, rng, 0.0, False, query=generic, chunk=20)       # 3 ragged chunks
        b = _render(net_c, net_f, ro, rd, rng, 0.0, False)
    report("sem_preds generic vs fused", a[4]["sem_preds"], b[4]["sem_preds"], atol=2e-2, rtol=1e-3)
    report("sem_preds0 generic vs fused", a[4]["sem_preds0"], b[4]["sem_preds0"], atol=2e-2, rtol=1e-3)
    report("rgb generic vs fused", a[0], b[0], atol=5e-3)
    assert a[4]["raw"].shape == b[4]["raw"].shape == (n, 128, 4 + K)


@pytest.mark.parametrize("ray_chunk", [1 << 20, 100])
def test_fused_train_step_with_semantic_targets_matches_drop_in_route(ray_chunk):
    n_rgb, n_dep = 192, 64
    lam, imp, slam = 0.01, 0.5, 0.05
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem = _case(n_rgb, n_dep, 73)
    d = dn()
    F = torch.nn.functional
    ts = tsem.to(DEV)
    rgb, disp, acc, depth, extras = _render(net_c, net_f, ro, rd, rng, 1.0, True)
    ce, ce0 = F.cross_entropy(extras["sem_preds"][:n_rgb], ts), F.cross_entropy(extras["sem_preds0"][:n_rgb], ts)
    loss = d.img2mse(rgb[:n_rgb], tgt.to(DEV)) + lam * imp * d.img2mse(depth[n_rgb:], dep.to(DEV)) \
        + slam * (ce + ce0) + d.img2mse(extras["rgb0"][:n_rgb], tgt.to(DEV))
    loss.backward()
    nets = list(net_c.named_parameters()) + list(net_f.named_parameters())
    ref = [q.grad.clone() for _, q in nets]
    for _, q in nets:
        q.grad = None
    inj = {k: getattr(rng, k).to(DEV) for k in ("t_rand", "noise0", "u", "noise1")}
    out = d.train_step(H, W, FOCAL, torch.stack([ro, rd], 0).to(DEV), tgt.to(DEV), dep.to(DEV), n_rgb, net_c, net_f,
                       N_samples=64, N_importance=64, perturb=1., raw_noise_std=1., depth_lambda=lam,
                       depth_importance=imp, target_semantic=ts, semantic_lambda=slam, ray_chunk=ray_chunk, _rng=inj)
    report("loss (fused vs drop-in)", out["loss"], loss, rtol=2e-5)
    report("semantic_loss", out["semantic_loss"], ce, rtol=2e-5)
    report("semantic_loss0", out["semantic_loss0"], ce0, rtol=2e-5)
    worst = max(rel_l2(q.grad, g) for (_, q), g in zip(nets, ref))
    print("  worst per-tensor rel-L2 between the two routes: %.3e" % worst)
    assert worst <= 2e-3


def test_graphed_train_step_with_semantic_targets():
    n_rgb, n_dep = 128, 64
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem = _case(n_rgb, n_dep, 79, False, False)
    d = dn()
    rays = torch.stack([ro, rd], 0).to(DEV)
    kw = dict(N_samples=64, N_importance=64, perturb=0., raw_noise_std=0., depth_lambda=0.01, depth_importance=1.,
              semantic_lambda=0.01)
    out = d.train_step(H, W, FOCAL, rays, tgt.to(DEV), dep.to(DEV), n_rgb, net_c, net_f, target_semantic=tsem.to(DEV), **kw)
    nets = list(net_c.parameters()) + list(net_f.parameters())
    ref, ref_loss = [q.grad.clone() for q in nets], out["loss"].item()
    step = d.GraphedTrainStep(H, W, FOCAL, n_rgb + n_dep, n_rgb, net_c, net_f, **kw)
    for rep in range(2):
        res = step(rays, tgt.to(DEV), dep.to(DEV), target_semantic=tsem.to(DEV))
        assert abs(res["loss"].item() - ref_loss) <= 1e-5 * abs(ref_loss)
        worst = max(rel_l2(q.grad, g) for q, g in zip(nets, ref))
        print("  replay %d: worst per-tensor rel-L2 vs eager %.3e" % (rep, worst))
        assert worst <= 2e-3


def test_render_path_returns_the_class_map(tmp_path):
    """render_path with semantic_loss in the kwargs returns a third array like the reference (run_nerf.py:355-357):
    here the arg-max class per pixel; no autograd graph, the head still reads the kept activations."""
    net_c, pc, spec_c, net_f, pf, spec_f, *_ = _case(8, 0, 83, False, False)
    d = dn()
    e_p, e_d = d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0]
    kw = dict(network_query_fn=d.FusedQuery(e_p, e_d, 1 << 16, 10, 4, 0), perturb=False, N_importance=64,
              network_fine=net_f, N_samples=64, network_fn=net_c, use_viewdirs=True, white_bkgd=False,
              raw_noise_std=0., ndc=True, semantic_loss=True)
    Hh, Ww, foc = 6, 10, 12.0
    pose = torch.eye(4, device=DEV)[:3, :4].clone()
    rgbs, disps, sems = d.render_path([pose], (Hh, Ww, foc), 32, kw)
    assert rgbs.shape == (1, Hh, Ww, 3) and sems.shape == (1, Hh, Ww) and sems.dtype == np.int64
    ro, rdw = d.get_rays(Hh, Ww, foc, pose)
    rb = O.pack_rays(Hh, Ww, foc, ro.reshape(-1, 3).cpu(), rdw.reshape(-1, 3).cpu())
    ref = O.render_rays(rb, pc, spec_c, pf, spec_f, 64, 64, O.RenderRNG(), semantic_loss=True)
    top2 = ref["sem_preds"].topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.2          # ignore pixels whose two best classes are within tolerance
    got = torch.from_numpy(sems.reshape(-1))
    assert torch.equal(got[clear], ref["sem_preds"].argmax(-1)[clear]) and int(clear.sum()) > 0


def test_fullsize_semantic_step_properties():
    """Config B with the head on (4096 rays, 64 + 64 samples, K = 19) through size-independent properties:
    the per-ray logits are the sum of the per-sample logits; a ray's logits do not depend on which other rays share
    the batch; the fused step's cross-entropy equals torch's on the rendered logits; doubling semantic_lambda with
    every other loss term switched off doubles every gradient (linearity through the fold / unfold kernels)."""
    n_rgb = n_dep = 2048
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, rng, tgt, dep, tsem = _case(n_rgb, n_dep, 91)
    d = dn()
    with torch.no_grad():
        rgb, disp, acc, depth, extras = _render(net_c, net_f, ro, rd, rng, 1.0, True)
    sp, sp0 = extras["sem_preds"], extras["sem_preds0"]
    report("sum of raw[..., 4:] == sem_preds (4096 rays)", extras["raw"][..., 4:].sum(1), sp, atol=5e-3, rtol=1e-4)
    idx = torch.arange(1000, 1100)
    sub = O.RenderRNG(**{k: getattr(rng, k)[idx] for k in ("t_rand", "noise0", "u", "noise1")})
    with torch.no_grad():
        part = _render(net_c, net_f, ro[idx], rd[idx], sub, 1.0, True, retraw=False)
    assert torch.equal(part[4]["sem_preds"], sp[idx.to(DEV)]) and torch.equal(part[4]["sem_preds0"], sp0[idx.to(DEV)])
    inj = {k: getattr(rng, k).to(DEV) for k in ("t_rand", "noise0", "u", "noise1")}
    rays, ts = torch.stack([ro, rd], 0).to(DEV), tsem.to(DEV)
    nets = list(net_c.parameters()) + list(net_f.parameters())

    def step(slam, lam=0., coarse=False):
        for q in nets:
            q.grad = None
        out = d.train_step(H, W, FOCAL, rays, tgt.to(DEV), dep.to(DEV), n_rgb, net_c, net_f, N_samples=64,
                           N_importance=64, perturb=1., raw_noise_std=1., depth_lambda=lam, target_semantic=ts,
                           semantic_lambda=slam, coarse_loss=coarse, _rng=inj)
        return out, [q.grad.clone() for q in nets]

    out, _ = step(0.01, lam=0.01, coarse=True)
    F = torch.nn.functional
    report("semantic_loss (4096 rays)", out["semantic_loss"], F.cross_entropy(sp[:n_rgb], ts), rtol=1e-5)
    report("semantic_loss0 (4096 rays)", out["semantic_loss0"], F.cross_entropy(sp0[:n_rgb], ts), rtol=1e-5)
    # the gradient is affine in semantic_lambda: g(lambda) = g_colour + lambda g_semantic
    _, g0 = step(1e-30)                      # colour share (fine net) through the same route; ~0 for the coarse net
    _, g1 = step(1.0)
    _, g3 = step(3.0)
    names = [n for net in (net_c, net_f) for n, _ in net.named_parameters()]
    worst = 0.0
    for n, a, b, c in zip(names, g1, g3, g0):
        if n.startswith(("views_linears", "rgb_linear", "alpha_linear")):
            continue                         # the colour / density heads see no semantic gradient (only atomics noise)
        da, db = a - c, b - c                # semantic share of the gradient at lambda = 1 and 3
        assert float(db.norm()) > 0, n
        worst = max(worst, rel_l2(3.0 * da, db))
    print("  worst rel-L2 of 3 x (grad at lambda) vs (grad at 3 lambda), semantic share: %.3e" % worst)
    assert worst <= 5e-3


@pytest.mark.parametrize("classes", [1, 2, 32])
def test_class_count_edges(classes):
    """K = 1, 2 and the maximum 32: per-point logits through NeRF.forward (warp-per-point kernel, keeps Hsum for the
    backward) and through forward_rays(point_logits=True) (tile-per-block packed-FMA kernel), per-ray logits as their
    sum, and the gradients of the head's own parameters."""
    net, p, spec = make_net(8, seed=40 + classes, semantic=classes)
    N, S = 5, 64                                                   # 320 points: two full tiles and a half-filled one
    ro, rd = O.synth_rays(N, seed=3, H=H, W=W, focal=FOCAL)
    rb = O.pack_rays(H, W, FOCAL, ro, rd)
    z = O.stratified_z(rb[:, 6:7], rb[:, 7:8], S, None)
    pts = rb[:, None, 0:3] + rb[:, None, 3:6] * z[:, :, None]
    x = torch.cat([O.posenc(pts.reshape(-1, 3), 10), O.posenc(rb[:, None, -3:].expand(N, S, 3).reshape(-1, 3), 4)], -1)
    p32 = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y32 = O.mlp_forward(p32, x, spec)
    y = net(x.to(DEV))
    assert y.shape == (N * S, 4 + classes)
    tol = dict(atol=2e-2, rtol=1e-2)
    report("K=%d forward(x) logits" % classes, y[:, 4:], y32[:, 4:], **tol)
    with torch.no_grad():
        raw, sem, logits = net.forward_rays(rb.to(DEV), z.to(DEV), semantic=True, point_logits=True)
    report("K=%d forward_rays point logits" % classes, logits.reshape(N * S, classes), y32[:, 4:], **tol)
    report("K=%d per-ray logits" % classes, sem, y32[:, 4:].reshape(N, S, classes).sum(1), atol=5e-2, rtol=1e-2)
    cot = torch.randn(N * S, classes, generator=torch.Generator().manual_seed(classes))
    (y[:, 4:] * cot.to(DEV)).sum().backward()
    (y32[:, 4:] * cot).sum().backward()
    for name in ("semantic_linear.1.weight", "semantic_linear.1.bias", "semantic_linear.0.weight", "feature_linear.bias"):
        g = dict(net.named_parameters())[name].grad
        assert rel_l2(g, p32[name].grad) <= 2e-2, (name, rel_l2(g, p32[name].grad))
