"""CPU tests: the C-ABI library loads and exports every symbol include/dlnerf_b200.h declares (no compute
calls without a GPU), the ctypes struct mirrors match the C structs, and the static MLP plan — executed
by the numpy interpreter in tests/plan_sim.py — reproduces the oracle MLP and its gradients."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import dlnerf_b200 as dn
from oracle import nerf_oracle as O
import plan_sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from importlib import import_module
plan_mod = import_module("depth-lidar-nerf_b200.plan")


@pytest.fixture(scope="module")
def lib():
    dn.build()
    return dn.lib()


def test_header_symbols_are_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "dlnerf_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(dln_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    bound = set(dn._lib.SIGNATURES) | {"dln_build_info"}
    assert declared == bound, (declared ^ bound)
    assert b"sm_100a" in lib.dln_build_info()


def test_struct_mirrors_match_the_c_structs(lib):
    out = (C.c_int * 6)()
    assert lib.dln_abi_sizes(out) == 0
    L = dn._lib
    assert list(out) == [C.sizeof(L.ChainStep), C.sizeof(L.ChainProgram), C.sizeof(L.ChainArgs),
                         C.sizeof(L.WgradItem), C.sizeof(L.PackJob), C.sizeof(L.SemOffsets)]


def test_invalid_arguments_return_minus_one_without_touching_the_gpu(lib):
    assert lib.dln_posenc(None, None, 4, 10, None) == -1
    assert lib.dln_searchsorted(None, 1, 1, None, 1, 1, None, 0, None) == -1
    assert lib.dln_composite_fwd(None, 4, None, None, None, 0.0, 0, None, None, None, None, None, 1, 64, None) == -1
    assert lib.dln_mlp_chain(None, None, 148, None) == -1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", dn._lib.SO_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing"
    assert "LDTM" in sass, "tcgen05.ld missing"
    assert "UBLKCP" in sass, "bulk (TMA engine) copies missing"
    assert "HMMA.16816" not in sass, "legacy mma.sync path present"


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "depth-lidar-nerf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            for line in src.splitlines():
                assert not re.match(r"\s*(from|import)\s+oracle", line), "%s imports the oracle: %s" % (fn, line)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    L = dn._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.lib()


def test_param_layout_matches_reference_order_and_counts():
    for D, n in ((8, 595844), (4, 316548)):
        shp = plan_mod.NetShape(D=D, input_ch=63, input_ch_views=27, use_viewdirs=True)
        names = [k for k, _ in shp.param_shapes()]
        assert names == list(O.MLPSpec(D=D).param_shapes().keys())
        assert sum(int(np.prod(s)) for _, s in shp.param_shapes()) == n
    net = dn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True)
    assert [k for k, _ in net.named_parameters()] == list(O.MLPSpec(D=8).param_shapes().keys())


def _flat_from(params, plan):
    flat = np.zeros(plan.n_params)
    for name, _ in plan.shape.param_shapes():
        v = params[name].double().numpy().reshape(-1)
        flat[plan.offsets[name]: plan.offsets[name] + v.size] = v
    return flat


@pytest.mark.parametrize("D,vd,fold,W", [(8, True, True, 256), (4, True, True, 256), (6, True, True, 256),
                                         (8, True, False, 256), (4, True, False, 256), (8, False, True, 256),
                                         (3, False, True, 256),
                                         # netwidth < 256 (run_nerf.py:693-700 takes any): zero-padded 256-column steps
                                         (8, True, True, 64), (8, True, True, 128), (4, True, True, 192),
                                         (8, False, True, 64), (6, False, True, 128)])
def test_plan_reproduces_oracle_forward_and_gradients(D, vd, fold, W):
    spec = O.MLPSpec(D=D, W=W, use_viewdirs=vd)
    params = {k: v.double() for k, v in O.init_params(spec, seed=D).items()}
    shape = plan_mod.NetShape(D=D, W=W, input_ch=63, input_ch_views=27, output_ch=5, use_viewdirs=vd)
    plan = plan_mod.build_plan(shape, fold_feature=fold)
    assert plan.fold == (fold and vd and W == 256)
    flat = plan_sim.extend_flat(plan, _flat_from(params, plan))      # what dln_mlp_fold appends (M, b')
    g = torch.Generator().manual_seed(D)
    P = 37
    x = torch.randn(P, 90, generator=g, dtype=torch.float64)
    cot = torch.randn(P, shape.out_ch, generator=g, dtype=torch.float64)
    pl = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y = O.mlp_forward(pl, x, spec)
    (y * cot).sum().backward()

    out, stash_f, masks = plan_sim.run_forward(plan, flat, x[:, :63].numpy(), x[:, 63:].numpy())
    np.testing.assert_allclose(out, y.detach().numpy(), atol=1e-10)
    stash_b = plan_sim.run_backward(plan, flat, cot.numpy(), masks)
    gflat = plan_sim.unfold_grads(plan, flat, plan_sim.run_wgrad(plan, stash_f, stash_b))
    for name, _ in shape.param_shapes():
        ref = pl[name].grad
        got = gflat[plan.offsets[name]: plan.offsets[name] + params[name].numel()].reshape(params[name].shape)
        if ref is None:
            assert np.all(got == 0), name
        else:
            np.testing.assert_allclose(got, ref.numpy(), atol=1e-9, err_msg=name)
    # slot bookkeeping: every slot below the advertised count is produced
    assert set(stash_f) == set(range(plan.fwd_slots)) - (set() if vd else {1})   # slot 1 = encoded direction
    assert set(stash_b) == set(range(plan.bwd_slots))
    assert plan.fwd.n_steps <= dn._lib.MAX_STEPS and plan.bwd.n_steps <= dn._lib.MAX_STEPS


def test_plan_rejects_what_the_kernels_cannot_do():
    with pytest.raises(NotImplementedError):      # widths are 64, 128, 192, 256 (zero-padded to the kernels' 256 columns)
        plan_mod.build_plan(plan_mod.NetShape(D=8, W=96, input_ch=63, input_ch_views=27))
    with pytest.raises(NotImplementedError):
        plan_mod.build_plan(plan_mod.NetShape(D=8, W=512, input_ch=63, input_ch_views=27))
    with pytest.raises(ValueError):
        plan_mod.build_plan(plan_mod.NetShape(D=5, input_ch=63, input_ch_views=27))   # skip after last layer
    with pytest.raises(NotImplementedError):     # the semantic kernels hold one class per register / lane: K <= 32
        dn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True, semantic_num_classes=33)._state()
