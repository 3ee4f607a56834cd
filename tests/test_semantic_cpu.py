"""CPU: the semantic head (run_nerf_helpers.py:107-111, :126-127, :586-593; run_nerf.py:1541-1548).

* the oracle against tests/golden/semantic.npz, which oracle/make_golden.py wrote from the UNMODIFIED reference
  (module forward + all gradients, raw2outputs(semantic_loss=True), full render + RGB / depth / cross-entropy loss
  + backward);
* the algebra csrc/semantic_kernels.cu implements (folded operands Sw / sc, per-ray activation sums, the input
  gradient G handed to the dgrad chain, the unfold formulas) restated in float64 numpy and checked against autograd
  through the oracle -- so a formula error shows up here, without a GPU;
* the plan's flat-buffer layout for the head.
"""
import os

import numpy as np
import torch

from oracle import nerf_oracle as O

T = torch.from_numpy
K = 19


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, "semantic.npz"))


def test_oracle_semantic_mlp_matches_reference(golden_dir):
    g = _load(golden_dir)
    spec = O.MLPSpec(D=8, W=64, semantic_num_classes=K)
    p = {k[6:]: T(g[k]).requires_grad_(True) for k in g.files if k.startswith("mlp_p_")}
    assert list(p) == list(spec.param_shapes())          # names AND registration order (optimizer / checkpoint order)
    y = O.mlp_forward(p, T(g["mlp_x"]), spec)
    assert y.shape[-1] == 4 + K
    np.testing.assert_allclose(y.detach().numpy(), g["mlp_y"], atol=2e-6, rtol=0)
    (y * T(g["mlp_cot"])).sum().backward()
    for k in g.files:
        if k.startswith("mlp_g_"):
            np.testing.assert_allclose(p[k[6:]].grad.numpy(), g[k], atol=2e-5, rtol=0)


def test_oracle_raw2outputs_semantic(golden_dir):
    g = _load(golden_dir)
    out = O.raw2outputs(T(g["r2o_raw"]), T(g["r2o_z"]), T(g["r2o_rays_d"]), None, False, semantic_loss=True)
    assert len(out) == 6
    np.testing.assert_allclose(out[5].numpy(), g["r2o_sem"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(out[0].numpy(), g["r2o_rgb"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(out[4].numpy(), g["r2o_depth"], atol=1e-6, rtol=0)


def test_oracle_semantic_render_loss_and_grads(golden_dir):
    g = _load(golden_dir)
    Hh, Ww, focal, n_rgb, n_dep = 94, 352, 138.14, 12, 8
    spec_c = O.MLPSpec(D=4, W=64, semantic_num_classes=K)
    spec_f = O.MLPSpec(D=8, W=64, semantic_num_classes=K)
    pc = {k: v.requires_grad_(True) for k, v in O.trained_like(O.init_params(spec_c, 201), 1.0).items()}
    pf = {k: v.requires_grad_(True) for k, v in O.trained_like(O.init_params(spec_f, 202)).items()}
    rb = O.pack_rays(Hh, Ww, focal, T(g["rays_o"]), T(g["rays_d"]), ndc=True, near=0.0, far=1.0, use_viewdirs=True)
    rng = O.RenderRNG(t_rand=T(g["t_rand"]), noise0=T(g["noise0"]), u=T(g["u"]), noise1=T(g["noise1"]))
    out = O.render_rays(rb, pc, spec_c, pf, spec_f, 64, 64, rng, raw_noise_std=1.0, semantic_loss=True)
    res = O.train_loss(out, n_rgb, T(g["tgt"]), T(g["dep"]), depth_lambda=0.01, depth_importance=0.5,
                       target_semantic=T(g["tsem"]), semantic_lambda=0.01)
    res["loss"].backward()
    np.testing.assert_allclose(out["sem_preds"].detach().numpy(), g["sem_preds"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(out["sem_preds0"].detach().numpy(), g["sem_preds0"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(out["rgb_map"].detach().numpy(), g["rgb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(res["loss"].item(), g["loss"], rtol=1e-6)
    for key, (params, name) in {"g_fine_sem1": (pf, "semantic_linear.1.weight"),
                                "g_fine_sem0": (pf, "semantic_linear.0.weight"),
                                "g_fine_feature_b": (pf, "feature_linear.bias"),
                                "g_fine_l7": (pf, "pts_linears.7.weight"),
                                "g_coarse_sem1_b": (pc, "semantic_linear.1.bias"),
                                "g_coarse_l0": (pc, "pts_linears.0.weight")}.items():
        np.testing.assert_allclose(params[name].grad.numpy(), g[key], atol=1e-5, rtol=0)


def test_folded_head_algebra_in_float64():
    """What csrc/semantic_kernels.cu computes, in numpy float64, against autograd through the layer-by-layer head."""
    rs = np.random.RandomState(3)
    W, Hd, N, S = 256, 128, 5, 7
    Wf, bf = rs.randn(W, W) / 16, rs.randn(W) / 4
    W1, b1 = rs.randn(Hd, W) / 16, rs.randn(Hd) / 4
    W2, b2 = rs.randn(K, Hd) / 11, rs.randn(K) / 4
    h = np.maximum(rs.randn(N, S, W), 0)                              # last trunk activations of N rays
    cot = rs.randn(N, K)                                              # d loss / d sem_preds

    # fold (sem_fold_a_kernel, sem_fold_s_kernel)
    A, a = W1 @ Wf, W1 @ bf + b1
    Sw, sc = W2 @ A, W2 @ a + b2
    # forward (sem_head_fwd_kernel)
    hsum = h.sum(1)
    sem = hsum @ Sw.T + S * sc
    # backward (sem_head_bwd_kernel): G is added to dH of every sample of the ray by the dgrad chain
    G = cot @ Sw
    dSw, dsc = cot.T @ hsum, S * cot.sum(0)
    # unfold (sem_unfold1_kernel, sem_unfold2_kernel)
    dW2, db2 = dSw @ A.T + np.outer(dsc, a), dsc
    dA, da = W2.T @ dSw, W2.T @ dsc
    dW1, db1 = dA @ Wf.T + np.outer(da, bf), da
    dWf, dbf = W1.T @ dA, W1.T @ da

    t = lambda x: torch.tensor(x, dtype=torch.float64, requires_grad=True)        # noqa: E731
    tWf, tbf, tW1, tb1, tW2, tb2, th = map(t, (Wf, bf, W1, b1, W2, b2, h))
    feat = th @ tWf.T + tbf
    logits = (feat @ tW1.T + tb1) @ tW2.T + tb2                                   # run_nerf_helpers.py:126-127
    ref = logits.sum(1)                                                           # :589
    np.testing.assert_allclose(sem, ref.detach().numpy(), rtol=1e-10, atol=1e-10)
    (ref * torch.tensor(cot)).sum().backward()
    for got, want in ((dW2, tW2), (db2, tb2), (dW1, tW1), (db1, tb1), (dWf, tWf), (dbf, tbf)):
        np.testing.assert_allclose(got, want.grad.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.broadcast_to(G[:, None, :], h.shape), th.grad.numpy(), rtol=1e-9, atol=1e-10)


def test_plan_layout_with_semantic_head():
    import dlnerf_b200 as d
    from dlnerf_b200 import plan as P
    shape = P.NetShape(D=8, semantic_num_classes=K)
    names = [n for n, _ in shape.param_shapes()]
    assert names[-4:] == ["semantic_linear.0.weight", "semantic_linear.0.bias", "semantic_linear.1.weight",
                          "semantic_linear.1.bias"]
    assert names == list(O.MLPSpec(D=8, semantic_num_classes=K).param_shapes())
    pl = P.build_plan(shape)
    so = pl.sem
    assert so.K == K and so.A == pl.off_bM + 128 and so.a == so.A + 128 * 256 and so.Sw == so.a + 128
    assert so.sc == so.Sw + K * 256 and pl.n_flat == so.sc + 20 and so.Sw % 4 == 0
    assert pl.h_last_slot == 2 + 4 * 7 and pl.h_last_slot + 4 <= pl.fwd_slots
    # the chain programs are those of the plain network: the head never enters the tensor-core chain
    plain = P.build_plan(P.NetShape(D=8))
    sig = lambda prog: [(st.n_out, st.nk, st.epi, st.stash_slot, st.mask_slot, st.n_heads)        # noqa: E731
                        for st in list(prog.steps)[:prog.n_steps]]
    assert sig(pl.fwd) == sig(plain.fwd) and sig(pl.bwd) == sig(plain.bwd)
    assert (pl.fwd_slots, pl.bwd_slots, len(pl.wgrad)) == (plain.fwd_slots, plain.bwd_slots, len(plain.wgrad))
    # without view directions the reference builds the layers but never evaluates them
    assert P.NetShape(D=8, use_viewdirs=False, semantic_num_classes=K).sem_K == 0
    assert P.build_plan(P.NetShape(D=8, use_viewdirs=False, semantic_num_classes=K)).sem is None
    # module: same parameter names / order as the reference's state_dict
    net = d.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True, semantic_num_classes=K)
    assert list(net.state_dict()) == names
