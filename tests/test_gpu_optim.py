"""FlatAdam (one dln_adam_step launch per network + bf16 re-pack) against torch.optim.Adam, the optimiser the
reference builds in create_nerf (run_nerf.py:440): same updates, same state layout, checkpoint interop, and the
reference's learning-rate decay by rewriting param_group['lr'] (run_nerf.py:1843-1847)."""
import copy

import pytest
import torch

from gpu_util import dn, make_net, report

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(D=4, seed=3):
    net, _, _ = make_net(D, seed=seed)
    twin = copy.deepcopy(net)
    return net, twin


def _set_grads(net, twin, seed, flat):
    g = torch.Generator(device=DEV).manual_seed(seed)
    if flat:    # the layout train_step produces: every .grad a view of one flat buffer in parameter order
        st = net._state()
        buf = torch.randn(net._plan.n_params, device=DEV, generator=g) * 1e-2
        for (name, _), p, q in zip(net._shape.param_shapes(), net._ordered_params(), twin._ordered_params()):
            o = net._plan.offsets[name]
            p.grad = buf[o:o + p.numel()].view(p.shape)
            q.grad = p.grad.clone()
    else:       # the layout autograd produces in general: separately allocated tensors
        for p, q in zip(net._ordered_params(), twin._ordered_params()):
            p.grad = torch.randn(p.shape, device=DEV, generator=g) * 1e-2
            q.grad = p.grad.clone()


@pytest.mark.parametrize("flat", [True, False])
def test_flat_adam_matches_torch_adam(flat):
    d = dn()
    net, twin = _pair()
    opt = d.FlatAdam([net], lr=5e-4, betas=(0.9, 0.999))
    ref = torch.optim.Adam(params=list(twin.parameters()), lr=5e-4, betas=(0.9, 0.999))
    for it in range(6):
        _set_grads(net, twin, 100 + it, flat)
        if it == 3:      # the reference's decay: rewrite lr in the param groups
            for o in (opt, ref):
                for gp in o.param_groups:
                    gp["lr"] = 5e-4 * (0.1 ** (it / 250000.0)) * 0.5
        opt.step()
        ref.step()
    for (n, p), q in zip(net.named_parameters(), twin.parameters()):
        report("param %s after 6 steps" % n, p, q, atol=1e-7, rtol=2e-6)
    for p, q in zip(net._ordered_params(), twin._ordered_params()):
        report("exp_avg", opt.state[p]["exp_avg"], ref.state[q]["exp_avg"], atol=1e-9, rtol=2e-6, quiet=True)
        report("exp_avg_sq", opt.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], atol=1e-12, rtol=2e-6, quiet=True)
        assert float(opt.state[p]["step"]) == float(ref.state[q]["step"]) == 6.0
    # the bf16 weight stages follow the update: the forward pass of `net` equals a fresh module holding twin's weights
    x = torch.randn(300, 90, device=DEV)
    fresh = d.NeRF(D=4, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).to(DEV)
    fresh.load_state_dict(twin.state_dict())
    with torch.no_grad():
        report("forward after FlatAdam steps", net(x), fresh(x), atol=2e-3)


def test_flat_adam_checkpoint_interop_with_torch_adam():
    """state_dict of one loads into the other (the reference reloads 'optimizer_state_dict', run_nerf.py:462-463)."""
    d = dn()
    net, twin = _pair(seed=5)
    opt = d.FlatAdam([net], lr=1e-3)
    ref = torch.optim.Adam(params=list(twin.parameters()), lr=1e-3, betas=(0.9, 0.999))
    for it in range(2):
        _set_grads(net, twin, 200 + it, True)
        opt.step()
        ref.step()
    # FlatAdam -> torch Adam and torch Adam -> FlatAdam, then one more step each way
    net2, twin2 = copy.deepcopy(net), copy.deepcopy(twin)
    ref2 = torch.optim.Adam(params=list(twin2.parameters()), lr=1e-3, betas=(0.9, 0.999))
    ref2.load_state_dict(opt.state_dict())
    opt2 = d.FlatAdam([net2], lr=1e-3)
    opt2.load_state_dict(ref.state_dict())
    _set_grads(net2, twin2, 300, False)
    opt2.step()
    ref2.step()
    for (n, p), q in zip(net2.named_parameters(), twin2.parameters()):
        report("param %s after reload + step" % n, p, q, atol=1e-7, rtol=2e-6, quiet=True)


def test_flat_adam_training_trajectory_matches_torch_adam():
    """8 iterations of train_step + optimiser on identical twins: FlatAdam (flat kernel + re-pack) and
    torch.optim.Adam (what create_nerf returns) must produce the same loss trajectory, and the loss must fall."""
    d = dn()
    from gpu_util import O
    nets = [make_net(4, seed=21, sigma_bias=1.0)[0], make_net(8, seed=22, sigma_bias=1.0)[0]]
    twins = [copy.deepcopy(n) for n in nets]
    opt = d.FlatAdam(nets, lr=2e-3)
    ref = torch.optim.Adam(params=[p for n in twins for p in n.parameters()], lr=2e-3, betas=(0.9, 0.999))
    ro, rd = O.synth_rays(512, seed=4)
    tgt, dep = O.synth_targets(256, 256, seed=4)
    rays = torch.stack([ro, rd], 0).to(DEV)
    rng = O.synth_rng(512, 64, 64, seed=4)
    inj = {k: getattr(rng, k).to(DEV) for k in ("t_rand", "noise0", "u", "noise1")}
    la, lb = [], []
    for it in range(8):
        for (nc, nf), o, acc in (((nets[0], nets[1]), opt, la), ((twins[0], twins[1]), ref, lb)):
            out = d.train_step(378, 504, 407.6, rays, tgt.to(DEV), dep.to(DEV), 256, nc, nf, depth_lambda=0.01, _rng=inj)
            acc.append(float(out["loss"]))
            o.step()
    print("  FlatAdam   %s" % " ".join("%.5f" % x for x in la))
    print("  torch Adam %s" % " ".join("%.5f" % x for x in lb))
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-3 * abs(y)
    assert la[-1] < la[0]


@pytest.mark.parametrize("route", [[], ["--drop-in"]])
def test_synthetic_scene_training_converges(route):
    """examples/train_synthetic.py: the reference's loop shape (ray loader -> render + RGB / depth loss -> backward ->
    Adam -> lr decay) on a synthetic view-dependent scene, through the graphed step and through the drop-in
    render() + loss.backward() route: PSNR must climb from ~11 dB to > 27 dB within 120 iterations."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_synthetic.py")
    spec = importlib.util.spec_from_file_location("train_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    log = mod.main(["--iters", "120", "--n-rand", "1024"] + route)
    print("  (loss, psnr) every 50 iterations:", log)
    assert log[0][1] < 15.0 and log[-1][1] > 27.0 and log[-1][0] < 0.1 * log[0][0]


@pytest.mark.parametrize("route", [[], ["--drop-in"]])
def test_synthetic_scene_training_with_semantic_head_converges(route):
    """The same loop with the semantic head on (4 classes = quadrant of the view direction, semantic_lambda 0.01 as
    in fern_dsnerf.txt:55-56): colour still converges and the cross-entropy of the fine per-ray logits falls well
    below its initial value -- the whole chain fold -> head -> cross-entropy -> dgrad row -> unfold -> Adam works."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_synthetic.py")
    spec = importlib.util.spec_from_file_location("train_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    log = mod.main(["--iters", "150", "--n-rand", "1024", "--semantic"] + route)
    print("  (loss, psnr, semantic CE) every 50 iterations:", log)
    assert log[-1][1] > 25.0
    assert log[-1][2] < 0.7 * log[0][2]
