"""CPU: the oracle (oracle/nerf_oracle.py) against the golden vectors that
oracle/make_golden.py produced by running the UNMODIFIED reference."""
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

T = torch.from_numpy


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_param_generator_has_not_drifted(golden_dir):
    g = load(golden_dir, "param_guard.npz")
    for D in (4, 8):
        pr = O.init_params(O.MLPSpec(D=D), seed=3407 + D)
        assert abs(float(sum(v.double().sum() for v in pr.values())) - g["D%d_sum" % D][0]) < 1e-6
        assert abs(float(sum(v.double().abs().sum() for v in pr.values())) - g["D%d_abs" % D][0]) < 1e-4
        np.testing.assert_array_equal(pr["pts_linears.0.weight"][:2, :8].numpy(), g["D%d_head" % D])
    # parameter counts quoted in SURVEY.md §8(a) R7
    assert sum(v.numel() for v in O.init_params(O.MLPSpec(D=8), 0).values()) == 595844
    assert sum(v.numel() for v in O.init_params(O.MLPSpec(D=4), 0).values()) == 316548


def test_posenc(golden_dir):
    g = load(golden_dir, "embed.npz")
    x = T(g["x"])
    np.testing.assert_array_equal(O.posenc(x, 10).numpy(), g["e10"])
    np.testing.assert_array_equal(O.posenc(x, 4).numpy(), g["e4"])
    assert O.posenc_dim(10) == 63 and O.posenc_dim(4) == 27 and O.posenc_dim(10, -1) == 3


@pytest.mark.parametrize("tag,D,vd", [("d8", 8, True), ("d4", 4, True), ("d8nv", 8, False)])
def test_mlp_forward_and_grads(golden_dir, tag, D, vd):
    g = load(golden_dir, "mlp_small.npz")
    spec = O.MLPSpec(D=D, W=64, use_viewdirs=vd)
    p = {k[len(tag) + 3:]: T(g[k]).requires_grad_(True) for k in g.files if k.startswith(tag + "_p_")}
    assert set(p) == set(spec.param_shapes())
    y = O.mlp_forward(p, T(g[tag + "_x"]), spec)
    np.testing.assert_allclose(y.detach().numpy(), g[tag + "_y"], atol=2e-6, rtol=0)
    (y * T(g[tag + "_cot"])).sum().backward()
    for k in g.files:
        if k.startswith(tag + "_g_"):
            np.testing.assert_allclose(p[k[len(tag) + 3:]].grad.numpy(), g[k], atol=2e-5, rtol=0)


@pytest.mark.parametrize("tag,std,wb", [("plain", 0.0, False), ("noise", 1.0, False), ("white", 0.5, True)])
def test_raw2outputs(golden_dir, tag, std, wb):
    g = load(golden_dir, "raw2outputs.npz")
    out = O.raw2outputs(T(g["raw"]), T(g["z"]), T(g["rays_d"]), T(g["noise"]) * std if std > 0 else None, wb)
    for name, a in zip(["rgb", "disp", "acc", "weights", "depth"], out):
        ref = g["%s_%s" % (tag, name)]
        ok = np.isfinite(ref)
        np.testing.assert_allclose(a.numpy()[ok], ref[ok], atol=1e-6 if name != "disp" else 0, rtol=1e-5)


def test_raw2outputs_backward(golden_dir):
    g = load(golden_dir, "raw2outputs.npz")
    raw = T(g["raw"]).requires_grad_(True)
    m = O.raw2outputs(raw, T(g["z"]), T(g["rays_d"]), T(g["noise"]), False)
    ((m[0] * T(g["c_rgb"])).sum() + (m[4] * T(g["c_dep"])).sum() + (m[2] * T(g["c_acc"])).sum()
     + (m[3] * T(g["c_w"])).sum()).backward()
    np.testing.assert_allclose(raw.grad.numpy(), g["draw"], atol=1e-5, rtol=1e-5)


def test_sample_pdf(golden_dir):
    g = load(golden_dir, "sample_pdf.npz")
    bins, w, u = T(g["bins"]), T(g["w"]), T(g["u"])
    np.testing.assert_array_equal(O.sample_pdf(bins, w, 64, u=u).numpy(), g["s_rand"])
    np.testing.assert_array_equal(O.sample_pdf(bins, w, 64, det=True).numpy(), g["s_det"])
    cdf = O.pdf_to_cdf(w)
    np.testing.assert_array_equal(cdf.numpy(), g["cdf"])
    np.testing.assert_array_equal(O.invert_cdf(bins, cdf, u)[1].numpy(), g["inds"])
    # indices are exactly numpy's side='right' on the same cdf (the vendored extension's contract)
    np.testing.assert_array_equal(O.searchsorted_rows(g["cdf"], g["u"], "right"), g["inds"])


def test_searchsorted_known_answer():
    # SURVEY §8(a) R9 probe: right=True on ties
    r = O.searchsorted_rows(np.array([[0, .25, .25, .5, 1.]], np.float32),
                            np.array([[0, .25, .3, .5, 1.]], np.float32), "right")
    assert r.tolist() == [[1, 3, 3, 4, 5]]


def test_searchsorted_grid(golden_dir):
    g = load(golden_dir, "searchsorted.npz")
    for i in range(int(g["n"][0])):
        side = "left" if g["side%d" % i][0] == 0 else "right"
        np.testing.assert_array_equal(O.searchsorted_rows(g["a%d" % i], g["v%d" % i], side), g["r%d" % i])


def _render_case(g):
    spec_c, spec_f = O.MLPSpec(D=4, W=64), O.MLPSpec(D=8, W=64)
    pc = O.trained_like(O.init_params(spec_c, 101), 1.0)
    pf = O.trained_like(O.init_params(spec_f, 102))
    rng = O.RenderRNG(T(g["t_rand"]), T(g["noise0"]), T(g["u"]), T(g["noise1"]))
    return spec_c, spec_f, pc, pf, rng


def test_ndc_and_render_and_loss(golden_dir):
    g = load(golden_dir, "render.npz")
    ro, rd = T(g["rays_o"]), T(g["rays_d"])
    o, d = O.ndc_rays(378, 504, 407.6, 1.0, ro, rd)
    np.testing.assert_allclose(o.numpy(), g["ndc_o"], atol=1e-6)
    np.testing.assert_allclose(d.numpy(), g["ndc_d"], atol=1e-6)
    assert "rgb" in g.files, "render fixture missing: make_golden.py could not import run_nerf"
    spec_c, spec_f, pc, pf, rng = _render_case(g)
    pcg = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    pfg = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    rb = O.pack_rays(378, 504, 407.6, ro, rd)
    out = O.render_rays(rb, pcg, spec_c, pfg, spec_f, 64, 64, rng, raw_noise_std=1.0)
    for mine, ref in [("rgb_map", "rgb"), ("depth_map", "depth"), ("acc_map", "acc"), ("rgb0", "rgb0"),
                      ("depth_map0", "depth0"), ("acc0", "acc0"), ("z_std", "z_std")]:
        np.testing.assert_allclose(out[mine].detach().numpy(), g[ref], atol=2e-6, rtol=0, err_msg=mine)
    np.testing.assert_allclose(out["raw"].detach().numpy(), g["raw"], atol=1e-5)
    n_rgb = g["tgt"].shape[0]
    res = O.train_loss(out, n_rgb, T(g["tgt"]), T(g["dep"]), depth_lambda=0.01, depth_importance=0.5)
    np.testing.assert_allclose(res["loss"].item(), float(g["loss"]), atol=1e-6)
    res["loss"].backward()
    np.testing.assert_allclose(pfg["pts_linears.0.weight"].grad.numpy(), g["g_fine_l0"], atol=1e-5)
    np.testing.assert_allclose(pfg["rgb_linear.weight"].grad.numpy(), g["g_fine_rgb"], atol=1e-5)
    np.testing.assert_allclose(pfg["alpha_linear.weight"].grad.numpy(), g["g_fine_alpha"], atol=1e-5)
    np.testing.assert_allclose(pcg["pts_linears.0.weight"].grad.numpy(), g["g_coarse_l0"], atol=1e-5)


def test_inverse_depth_smoothness_golden(golden_dir):
    """oracle.inverse_depth_smoothness against InverseDepthSmoothnessLoss of the reference (loss.py:55-133): value
    and both gradients, bit-exact (same torch ops in the same order)."""
    g = np.load(os.path.join(golden_dir, "inv_depth_smooth.npz"))
    d = torch.from_numpy(g["idepth"]).requires_grad_(True)
    im = torch.from_numpy(g["image"]).requires_grad_(True)
    loss = O.inverse_depth_smoothness(d, im)
    loss.backward()
    assert float(loss.detach()) == float(g["loss"])
    assert np.array_equal(d.grad.numpy(), g["g_idepth"]) and np.array_equal(im.grad.numpy(), g["g_image"])


def test_ray_generators(golden_dir):
    """get_rays_np / get_rays_by_coord_np / get_rays_cropped_feature_loss_new, bit for bit."""
    g = load(golden_dir, "raygen.npz")
    H, W, focal = int(g["HWf"][0]), int(g["HWf"][1]), float(g["HWf"][2])
    for n in range(3):
        o, d = O.get_rays_np(H, W, focal, g["poses"][n])
        np.testing.assert_array_equal(d, g["grid_d%d" % n])
        np.testing.assert_array_equal(o, np.broadcast_to(g["poses"][n][:, 3], d.shape))
    for tag in ("64", "32"):
        o, d = O.get_rays_by_coord_np(H, W, focal, g["poses"][1], g["coord" + tag])
        assert d.dtype == g["coord_d" + tag].dtype
        np.testing.assert_array_equal(d, g["coord_d" + tag])
        np.testing.assert_array_equal(o, g["coord_o" + tag])
    for n in range(int(g["n_crops"][0])):
        nH, nW, gH, gW, sw, sh = (int(v) for v in g["crop%d_cfg" % n])
        grad, nograd, crop = O.rays_cropped_feature_loss_new(H, W, focal, g["poses"][2], nH, nW, gH, gW, sw, sh,
                                                             g["crop%d_perm" % n])
        assert crop == [sw, sw + nW - 1, sh, sh + nH - 1]
        np.testing.assert_array_equal(np.concatenate([grad[1], nograd[1]], 0), g["crop%d_d" % n])
        np.testing.assert_array_equal(np.concatenate([grad[2], nograd[2]], 0), g["crop%d_pts" % n])
        assert grad[0].shape == (gH * gW, 3) and nograd[0].shape == (nH * nW - gH * gW, 3)


@pytest.mark.parametrize("D,fname", [(8, "mlp_w256.npz"), (4, "mlp_w256_d4.npz")])
def test_mlp_w256_reference_case(golden_dir, D, fname):
    """The full-width cases generated by the unmodified reference module (forward + every parameter gradient)."""
    g = load(golden_dir, fname)
    spec = O.MLPSpec(D=D)
    params = O.trained_like(O.init_params(spec, seed=int(g["seed"][0])), float(g["sigma_bias"][0]))
    pl = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y = O.mlp_forward(pl, T(g["x"]), spec)
    np.testing.assert_allclose(y.detach().numpy(), g["y"], atol=5e-6, rtol=0)
    (y * T(g["cot"])).sum().backward()
    for k, v in pl.items():
        gr = v.grad.numpy()
        ref = g["g_" + k]
        got = gr[::16] if gr.ndim == 2 and gr.shape[0] >= 128 else gr
        np.testing.assert_allclose(got, ref, atol=1e-4, rtol=0)
        assert abs(float(v.grad.double().norm()) - g["gn_" + k][0]) <= 1e-4 * max(1.0, g["gn_" + k][0])
