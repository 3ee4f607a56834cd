"""GPU tests of the in-kernel random numbers (include/dlnerf_b200.h, "In-kernel random numbers") and of the routes that
run with LIVE randomness: the reference's four per-render draws (run_nerf.py:585, run_nerf_helpers.py:509, :565 twice)
are generated inside the consuming kernels.  `dln_rng_fill` writes the very numbers a kernel draws, so every live
route is checked against the injected-tensor route (and through it against the oracle) on identical values -- not
statistically."""
import math

import pytest
import torch

from gpu_util import O, dn, make_net, rel_l2, report

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, W, FOCAL = 378, 504, 407.6


def _state(seed=1234):
    return dn().ops.RngState(DEV, seed)


def test_draw_moments_and_independence():
    """Uniforms lie in [0, 1) on the 24-bit grid with the moments of U[0,1); normals have the moments of N(0,1);
    tensors with different names (offsets), different seeds, or after `advance` are different and uncorrelated."""
    st = _state()
    n = 1 << 20
    u = st.fill(0, "u", n // 64, 64).double()
    assert u.min().item() >= 0.0 and u.max().item() < 1.0
    assert torch.equal(u * (1 << 24), torch.round(u * (1 << 24)))
    se = 1.0 / math.sqrt(12 * n)
    assert abs(u.mean().item() - 0.5) < 5 * se
    assert abs(u.var().item() - 1 / 12) < 5 * math.sqrt(1 / 180 / n)
    hist = torch.histc(u.float(), 64, 0, 1)
    chi2 = ((hist - n / 64) ** 2 / (n / 64)).sum().item()
    assert chi2 < 63 + 6 * math.sqrt(2 * 63), chi2                 # chi-square with 63 dof
    g = st.fill(1, "n", n // 128, 128).double()
    assert abs(g.mean().item()) < 5 / math.sqrt(n)
    assert abs(g.var().item() - 1) < 5 * math.sqrt(2 / n)
    assert abs((g ** 3).mean().item()) < 5 * math.sqrt(15 / n)
    assert abs((g ** 4).mean().item() - 3) < 5 * math.sqrt(96 / n)
    assert g.abs().max().item() < 6.0
    # lag-1 correlation inside a tensor (neighbouring elements share a Philox block / a Box-Muller pair)
    gf = g.flatten()
    assert abs((gf[:-1] * gf[1:]).mean().item()) < 5 / math.sqrt(n)
    uf = u.flatten() - 0.5
    assert abs((uf[:-1] * uf[1:]).mean().item()) < 5 / 12 / math.sqrt(n)
    # names
    u1 = st.fill(1, "u", n // 64, 64).double()
    assert abs(((u - 0.5) * (u1 - 0.5)).mean().item()) < 5 / 12 / math.sqrt(n)
    assert torch.equal(st.fill(0, "u", 8, 64).double(), u[:8])                      # reproducible
    assert not torch.equal(_state(99).fill(0, "u", 8, 64).double(), u[:8])          # seed
    st.advance(1)
    assert torch.equal(st.fill(0, "u", n // 64, 64).double(), u1)                   # base + offset names the tensor
    ur = st.fill(5, "u_resample", 4096, 64).double()
    assert abs(ur.mean().item() - 0.5) < 5 / math.sqrt(12 * ur.numel())


@pytest.mark.parametrize("N,S,Ni", [(300, 64, 64), (77, 48, 40), (65, 96, 128)])
def test_kernels_draw_what_rng_fill_reports(N, S, Ni):
    """Each `_rng` kernel equals its tensor-fed twin bit for bit when the twin is fed dln_rng_fill's output."""
    d = dn()
    st = _state(7)
    ro, rd = O.synth_rays(N, seed=3)
    rb = d.ops.pack_rays(H, W, FOCAL, ro.to(DEV), rd.to(DEV), True, 0., 1., True)
    z_live = d.ops.stratified_z(rb, S, rng=(st, 3))
    z_inj = d.ops.stratified_z(rb, S, st.fill(3, "u", N, S))
    assert torch.equal(z_live, z_inj)
    raw = torch.randn(N, S, 4, device=DEV, generator=torch.Generator(DEV).manual_seed(5))
    raw.requires_grad_(True)
    rays_d = rb[:, 3:6].contiguous()
    live = d.ops.composite(raw, z_live, rays_d, None, 1.0, False, rng=(st, 4))
    inj = d.ops.composite(raw, z_live, rays_d, st.fill(4, "n", N, S), 1.0, False)
    for a, b in zip(live, inj):
        assert torch.equal(a, b)
    ga, = torch.autograd.grad(live[0].sum() + live[4].sum() + (live[3] ** 2).sum(), raw)
    gb, = torch.autograd.grad(inj[0].sum() + inj[4].sum() + (inj[3] ** 2).sum(), raw)
    assert torch.equal(ga, gb)                                  # the backward regenerates the same noise
    w = live[3].detach()
    fast = S <= 64 and Ni <= 64
    zs_live, zm_live = d.ops.importance_resample(z_live, w, Ni, rng=(st, 9))
    zs_inj, zm_inj = d.ops.importance_resample(z_live, w, Ni, st.fill(9, "u_resample" if fast else "u", N, Ni))
    assert torch.equal(zm_live, zm_inj) and torch.equal(zs_live, zs_inj)
    bins = torch.sort(torch.rand(N, S, device=DEV), -1)[0]
    a = d.ops.sample_pdf(bins, w[:, :-1].contiguous(), Ni, rng=(st, 11))
    b = d.ops.sample_pdf(bins, w[:, :-1].contiguous(), Ni, st.fill(11, "u", N, Ni))
    assert torch.equal(a, b)


def _case(n_rgb, n_dep, seed):
    net_c, pc, spec_c = make_net(4, seed=seed)
    net_f, pf, spec_f = make_net(8, seed=seed + 1)
    ro, rd = O.synth_rays(n_rgb + n_dep, seed=seed)
    tgt, dep = O.synth_targets(n_rgb, n_dep, seed=seed)
    return net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, tgt, dep


def _fills(st, N, base=0):
    return {"t_rand": st.fill(base + 0, "u", N, 64), "noise0": st.fill(base + 1, "n", N, 64),
            "u": st.fill(base + 2, "u_resample", N, 64), "noise1": st.fill(base + 3, "n", N, 128)}


def test_live_train_step_equals_injected_and_oracle():
    """train_step with in-kernel draws == train_step fed the same draws as tensors (same kernels: 2e-3 for the fp32
    atomics), and its loss matches the fp32 oracle evaluated on those draws (rel 2e-2, the e2e tolerance)."""
    n_rgb, n_dep = 160, 96
    N = n_rgb + n_dep
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, tgt, dep = _case(n_rgb, n_dep, 61)
    d = dn()
    st = _state(2024)
    rays = torch.stack([ro, rd], 0).to(DEV)
    kw = dict(N_samples=64, N_importance=64, perturb=1., raw_noise_std=1., depth_lambda=0.01, depth_importance=0.5)
    inj = _fills(st, N)
    nets = list(net_c.parameters()) + list(net_f.parameters())
    out_l = d.train_step(H, W, FOCAL, rays, tgt.to(DEV), dep.to(DEV), n_rgb, net_c, net_f, rng_state=st, **kw)
    g_live = [p.grad.clone() for p in nets]
    assert int(st.state[1].item()) == 4                      # the step advanced the generator
    out_i = d.train_step(H, W, FOCAL, rays, tgt.to(DEV), dep.to(DEV), n_rgb, net_c, net_f, _rng=inj, **kw)
    for k in ("loss", "img_loss", "img_loss0", "depth_loss"):
        report("live vs injected %s" % k, out_l[k], out_i[k], rtol=1e-5)
    worst = max(rel_l2(p.grad, g) for p, g in zip(nets, g_live))
    print("  worst per-tensor rel-L2 live vs injected: %.3e" % worst)
    assert worst <= 2e-3
    # oracle on the same numbers
    rng = O.RenderRNG(t_rand=inj["t_rand"].cpu(), noise0=inj["noise0"].cpu(), u=inj["u"].cpu(), noise1=inj["noise1"].cpu())
    rb = O.pack_rays(H, W, FOCAL, ro, rd)
    ref = O.render_rays(rb, pc, spec_c, pf, spec_f, 64, 64, rng, raw_noise_std=1.0)
    res = O.train_loss(ref, n_rgb, tgt, dep, depth_lambda=0.01, depth_importance=0.5)
    report("live loss vs oracle", out_l["loss"], res["loss"], rtol=2e-2)
    # a second live step names fresh tensors
    out_2 = d.train_step(H, W, FOCAL, rays, tgt.to(DEV), dep.to(DEV), n_rgb, net_c, net_f, rng_state=st, **kw)
    assert out_2["loss"].item() != out_l["loss"].item()


def test_graphed_step_live_randomness_schedule_and_depth_norm():
    """The route the bench's `value` runs (GraphedTrainStep, perturb = 1, raw_noise_std = 1): every replay draws
    fresh numbers, and replay k equals the eager step fed -- as tensors -- the draws the graph's generator names at
    that moment; the depth_importance decay (run_nerf.py:1527-1532) and the 'weighted_norm' depth mode (:1518) are
    followed without re-capturing."""
    n_rgb, n_dep = 128, 128
    N = n_rgb + n_dep
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, tgt, dep = _case(n_rgb, n_dep, 71)
    d = dn()
    rays = torch.stack([ro, rd], 0).to(DEV)
    rw = (0.5 + torch.rand(n_dep, generator=torch.Generator().manual_seed(3))).to(DEV)
    kw = dict(N_samples=64, N_importance=64, perturb=1., raw_noise_std=1., depth_lambda=0.05,
              depth_mode="weighted_norm")
    st = _state(77)
    step = d.GraphedTrainStep(H, W, FOCAL, N, n_rgb, net_c, net_f, rng_state=st, use_ray_weights=True,
                              depth_importance=1.0, **kw)
    nets = list(net_c.parameters()) + list(net_f.parameters())
    losses = []
    for rep, imp in enumerate([1.0, 0.5, 0.25]):
        base = int(st.state[1].item())
        assert base == 4 * (rep + 3)                  # 3 warm-up passes (capturing runs nothing) + `rep` replays
        inj = _fills(st, N)                           # the tensors the coming replay will draw
        dscale = 1.0 + rep                            # changes max(target_depth) between replays
        res = step(rays, tgt.to(DEV), dep.to(DEV) * dscale, ray_weights=rw, depth_importance=imp)
        got = {k: v.item() for k, v in res.items()}
        g_graph = [p.grad.clone() for p in nets]
        ref = d.train_step(H, W, FOCAL, rays, tgt.to(DEV), dep.to(DEV) * dscale, n_rgb, net_c, net_f, ray_weights=rw,
                           depth_importance=imp, _rng=inj, **kw)
        for k in ("loss", "img_loss", "img_loss0", "depth_loss"):
            report("replay %d %s" % (rep, k), got[k], ref[k], rtol=1e-5)
        worst = max(rel_l2(g, p.grad) for g, p in zip(g_graph, nets))
        print("  replay %d: worst per-tensor rel-L2 graph vs eager-injected %.3e" % (rep, worst))
        assert worst <= 2e-3
        for p, g in zip(nets, g_graph):
            p.grad = g
        losses.append(got["img_loss"])
    assert len(set(losses)) == 3                      # fresh draws on every replay


def test_drop_in_render_live_randomness_matches_oracle():
    """render() + loss.backward() with in-kernel draws against the fp32 oracle fed the same numbers."""
    n_rgb, n_dep = 96, 32
    N = n_rgb + n_dep
    net_c, pc, spec_c, net_f, pf, spec_f, ro, rd, tgt, dep = _case(n_rgb, n_dep, 81)
    d = dn()
    st = _state(5)
    off = st.calls
    q = d.FusedQuery(*d.get_embedder(10, 0)[:1], d.get_embedder(4, 0)[0], 1 << 16, 10, 4, 0)
    rgb, disp, acc, depth, extras = d.render(
        H, W, FOCAL, chunk=1 << 20, rays=torch.stack([ro, rd], 0).to(DEV), near=0., far=1., network_query_fn=q,
        perturb=1.0, N_importance=64, network_fine=net_f, N_samples=64, network_fn=net_c, use_viewdirs=True,
        white_bkgd=False, raw_noise_std=1.0, ndc=True, _rng={"state": st})
    assert st.calls == off + 4
    inj = _fills(st, N, off)
    rng = O.RenderRNG(t_rand=inj["t_rand"].cpu(), noise0=inj["noise0"].cpu(), u=inj["u"].cpu(), noise1=inj["noise1"].cpu())
    pcg = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    pfg = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    ref = O.render_rays(O.pack_rays(H, W, FOCAL, ro, rd), pcg, spec_c, pfg, spec_f, 64, 64, rng, raw_noise_std=1.0)
    report("rgb_map", rgb, ref["rgb_map"], atol=2e-2)
    report("depth_map", depth, ref["depth_map"], atol=2e-2)
    report("rgb0", extras["rgb0"], ref["rgb0"], atol=2e-2)
    res = O.train_loss(ref, n_rgb, tgt, dep, depth_lambda=0.01, depth_importance=1.0)
    res["loss"].backward()
    loss = d.img2mse(rgb[:n_rgb], tgt.to(DEV)) + 0.01 * d.img2mse(depth[n_rgb:], dep.to(DEV)) + \
        d.img2mse(extras["rgb0"][:n_rgb], tgt.to(DEV))
    report("loss", loss, res["loss"], rtol=2e-2)
    loss.backward()                  # the compositing backward regenerates its noise from the tensor names
    num = sum(((p.grad.cpu().double() - pfg[n].grad.double()) ** 2).sum() for n, p in net_f.named_parameters())
    den = sum((pfg[n].grad.double() ** 2).sum() for n, p in net_f.named_parameters())
    agg = math.sqrt(num / den)
    print("  fine-net aggregate gradient rel-L2 vs fp32 oracle: %.3e" % agg)
    assert agg <= 2e-2


def test_no_grad_forward_writes_no_stash():
    """under torch.no_grad() the forward-only kernels run -- no activation stash, no ReLU masks."""
    net, _, _ = make_net(8, seed=3)
    d = dn()
    N, S = 2048, 128
    ro, rd = O.synth_rays(N, seed=3)
    rb = d.ops.pack_rays(H, W, FOCAL, ro.to(DEV), rd.to(DEV), True, 0., 1., True)
    z = d.ops.stratified_z(rb, S)
    with torch.no_grad():
        net.forward_rays(rb, z)                      # plans, weight packs
    torch.cuda.synchronize()

    def peak(fn):
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        out = fn()
        torch.cuda.synchronize()
        p = torch.cuda.max_memory_allocated() - base
        del out
        return p

    def nograd():
        with torch.no_grad():
            return net.forward_rays(rb, z)

    p_inf, p_train = peak(nograd), peak(lambda: net.forward_rays(rb, z))
    stash = N * S * 4608                             # 36 slabs x 16 KB per 128 points
    print("  peak bytes: no_grad %.1f MB, grad %.1f MB (stash %.1f MB)" % (p_inf / 1e6, p_train / 1e6, stash / 1e6))
    assert p_inf < 0.05 * stash and p_train >= stash


@pytest.mark.parametrize("perturb,lindisp", [(True, False), (False, False), (True, True)])
def test_fused_stratified_sampling_gives_the_stand_alone_kernels_bits(perturb, lindisp):
    """north_star part 1: the coarse chain's tile prologue computes the stratified depths it encodes (run_nerf.py:571-593)
    and writes z_vals; they must equal dln_stratified_z(_rng) bit for bit (same draws, same arithmetic), and so must the
    network output that was computed from them."""
    d = dn()
    net, _, _ = make_net(4, seed=5)
    N, S = 300, 64                       # 150 tiles: several per CTA pair, the last one partial
    ro, rd = O.synth_rays(N, seed=12)
    rb = d.pack_ray_batch(378, 504, 407.6, torch.stack([ro, rd], 0).to(DEV), ndc=not lindisp, near=0.5 if lindisp else 0.,
                          far=6. if lindisp else 1.)
    st = d.ops.RngState(torch.device(DEV), 1234)
    rng = (st, 7) if perturb else None
    z_ref = d.ops.stratified_z(rb, S, None, lindisp, rng=rng)
    with torch.no_grad():
        raw_ref = net.forward_rays(rb, z_ref)
        z = torch.full((N, S), float("nan"), device=DEV)
        raw = net.forward_rays(rb, z, strat=dict(rng=rng, lindisp=lindisp))
    assert torch.equal(z, z_ref), "fused stratified depths differ from the stand-alone kernel"
    assert torch.equal(raw, raw_ref)
