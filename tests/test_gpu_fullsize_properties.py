"""Config B at its full size (4096 rays: 2048 RGB + 2048 depth, 64 + 64 samples, D=4 coarse / D=8 fine) checked
through properties that do not need the CPU oracle to finish at that size:

* hierarchical sampling: merged z values are ascending, are exactly the multiset {coarse z} U {new samples},
  the indices are bit-exact ``searchsorted(cdf, u, right=True)`` on the kernel's own cdf;
* compositing: weights >= 0, acc_map == sum(weights) <= 1 (+ rounding), depth within [near, far];
* forward determinism: two evaluations of the same batch are bit-identical (no atomics on the forward path);
* gradient linearity: doubling every loss coefficient doubles every parameter gradient (fp32 atomics only
  reorder sums: rel-L2 <= 1e-3);
* ray-chunking, graph replay and the drop-in autograd route agree with the one-shot fused step;
* a sampled subset of 64 rays agrees with the oracle (the pinned restatement of the reference) within the
  bf16 tolerances stated in test_gpu_render_e2e.py.
"""
import pytest
import torch

from gpu_util import O, dn, make_net, rel_l2, report

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, W, FOCAL = 378, 504, 407.6
N_RGB = N_DEP = 2048


@pytest.fixture(scope="module")
def case():
    net_c, pc, spec_c = make_net(4, seed=7, sigma_bias=1.0)
    net_f, pf, spec_f = make_net(8, seed=8, sigma_bias=1.0)
    ro, rd = O.synth_rays(N_RGB + N_DEP, seed=7)
    rng = O.synth_rng(N_RGB + N_DEP, 64, 64, seed=7, perturb=True, noise=True)
    tgt, dep = O.synth_targets(N_RGB, N_DEP, seed=7)
    inj = {k: getattr(rng, k).to(DEV) for k in ("t_rand", "noise0", "u", "noise1")}
    return dict(net_c=net_c, net_f=net_f, pc=pc, pf=pf, spec_c=spec_c, spec_f=spec_f, ro=ro, rd=rd, rng=rng,
                rays=torch.stack([ro, rd], 0).to(DEV), tgt=tgt.to(DEV), dep=dep.to(DEV), inj=inj)


def _render(c):
    d = dn()
    q = d.FusedQuery(d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0], 1 << 16, 10, 4, 0)
    kw = dict(network_query_fn=q, perturb=1.0, N_importance=64, network_fine=c["net_f"], N_samples=64,
              network_fn=c["net_c"], use_viewdirs=True, white_bkgd=False, raw_noise_std=1.0, ndc=True,
              _rng=c["inj"])
    with torch.no_grad():
        return d.render(H, W, FOCAL, chunk=1 << 20, rays=c["rays"], retraw=True, near=0., far=1., **kw)


def test_fullsize_sampling_and_compositing_properties(case):
    d = dn()
    rb = d.pack_ray_batch(H, W, FOCAL, case["rays"], True, 0., 1., True)
    z0 = d.ops.stratified_z(rb, 64, case["inj"]["t_rand"])
    assert bool((z0[:, 1:] >= z0[:, :-1]).all()) and float(z0.min()) >= 0. and float(z0.max()) <= 1.
    raw0 = case["net_c"].forward_rays(rb, z0).detach()
    rgb0, disp0, acc0, w0, depth0 = d.ops.composite(raw0, z0, rb[:, 3:6].contiguous(), case["inj"]["noise0"], 1.0, False)
    assert float(w0.min()) >= 0.
    report("acc == sum(weights)", acc0, w0.sum(-1), atol=2e-6, rtol=1e-5)
    assert float(acc0.max()) <= 1. + 1e-5
    ok = acc0 > 1e-3
    q = (depth0 / acc0)[ok]
    assert float(q.min()) >= -1e-5 and float(q.max()) <= 1. + 1e-5        # a convex combination of z in [0, 1]
    zs, zm, cdf, inds = d.ops.importance_resample(z0, w0, 64, case["inj"]["u"], return_debug=True)
    assert bool((zm[:, 1:] >= zm[:, :-1]).all()), "merged z values must ascend"
    both = torch.sort(torch.cat([z0, zs], -1), -1)[0]
    assert torch.equal(zm, both), "merged z values must be exactly sort(cat(z_vals, z_samples))"
    ref_inds = torch.searchsorted(cdf.contiguous(), case["inj"]["u"].contiguous(), right=True)
    assert torch.equal(inds, ref_inds), "sample_pdf indices must be bit-exact on the kernel's own cdf"


def test_fullsize_forward_is_deterministic(case):
    a, b = _render(case), _render(case)
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    assert torch.equal(a[4]["raw"], b[4]["raw"])


def _step(case, scale=1.0, **extra):
    d = dn()
    nets = list(case["net_c"].parameters()) + list(case["net_f"].parameters())
    for p in nets:
        p.grad = None
    # depth_importance scales only the depth term; scale the colour terms by scaling the targets' loss through
    # a wrapper is not possible, so linearity is tested on the depth term (coef_dep) with the colour loss fixed
    out = d.train_step(H, W, FOCAL, case["rays"], case["tgt"], case["dep"], N_RGB, case["net_c"], case["net_f"],
                       N_samples=64, N_importance=64, perturb=1., raw_noise_std=1., depth_lambda=0.01,
                       depth_importance=scale, _rng=case["inj"], **extra)
    return out, [p.grad.clone() for p in nets]


def test_fullsize_gradient_is_affine_in_the_depth_weight(case):
    """grad(depth_importance = s) = g_rgb + s * g_depth, so g(0), g(1), g(3) must be collinear per tensor."""
    _, g0 = _step(case, 0.0)
    _, g1 = _step(case, 1.0)
    _, g3 = _step(case, 3.0)
    worst = 0.0
    for a, b, c in zip(g0, g1, g3):
        pred = a + 3.0 * (b - a)
        worst = max(worst, rel_l2(c, pred))
    print("  worst per-tensor rel-L2 of g(3) against g(0) + 3 (g(1) - g(0)): %.3e" % worst)
    assert worst <= 2e-3


def test_fullsize_routes_agree(case):
    d = dn()
    ref_out, ref = _step(case, 1.0, ray_chunk=1 << 20)
    out_c, g_c = _step(case, 1.0, ray_chunk=1000)
    report("chunked loss", out_c["loss"], ref_out["loss"], rtol=2e-5)
    worst = max(rel_l2(a, b) for a, b in zip(g_c, ref))
    print("  ray_chunk=1000 vs one shot: worst per-tensor rel-L2 %.3e" % worst)
    assert worst <= 2e-3
    # drop-in autograd route on the same injected draws
    nets = list(case["net_c"].parameters()) + list(case["net_f"].parameters())
    for p in nets:
        p.grad = None
    q = d.FusedQuery(d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0], 1 << 16, 10, 4, 0)
    rgb, disp, acc, depth, extras = d.render(
        H, W, FOCAL, chunk=1 << 20, rays=case["rays"], retraw=True, near=0., far=1., network_query_fn=q, perturb=1.0,
        N_importance=64, network_fine=case["net_f"], N_samples=64, network_fn=case["net_c"], use_viewdirs=True,
        white_bkgd=False, raw_noise_std=1.0, ndc=True, _rng=case["inj"])
    loss = d.img2mse(rgb[:N_RGB], case["tgt"]) + 0.01 * d.img2mse(depth[N_RGB:], case["dep"]) \
        + d.img2mse(extras["rgb0"][:N_RGB], case["tgt"])
    loss.backward()
    report("drop-in loss", loss, ref_out["loss"], rtol=2e-5)
    worst = max(rel_l2(p.grad, b) for p, b in zip(nets, ref))
    print("  drop-in autograd route vs fused step: worst per-tensor rel-L2 %.3e" % worst)
    assert worst <= 2e-3


def test_fullsize_subset_against_oracle(case):
    """64 of the 4096 rays through the oracle (rays are independent, so a subset of the batch is a valid check)."""
    out = _render(case)
    idx = torch.arange(0, N_RGB + N_DEP, 64)
    rb = O.pack_rays(H, W, FOCAL, case["ro"][idx], case["rd"][idx])
    rng = case["rng"]
    sub = O.RenderRNG(*(getattr(rng, k)[idx] for k in ("t_rand", "noise0", "u", "noise1")))
    ref = O.render_rays(rb, case["pc"], case["spec_c"], case["pf"], case["spec_f"], 64, 64, sub, raw_noise_std=1.0)
    report("rgb_map (64-ray subset)", out[0][idx.to(DEV)], ref["rgb_map"], atol=2e-2)
    report("depth_map (64-ray subset)", out[3][idx.to(DEV)], ref["depth_map"], atol=2e-2)
    report("acc_map (64-ray subset)", out[2][idx.to(DEV)], ref["acc_map"], atol=2e-2)
