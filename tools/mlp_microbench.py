"""GPU microbenchmark of the MLP kernels in isolation (CUDA events, 10 reps after 3 warm-ups)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
L = dn._lib
DEV = "cuda"
N, S = int(os.environ.get("NRAYS", 4096)), 128
torch.manual_seed(0)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for D in (8, 4):
    net = dn.NeRF(D=D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(DEV)
    rb = torch.randn(N, 11, device=DEV); rb[:, 6] = 0; rb[:, 7] = 1
    z = torch.sort(torch.rand(N, S, device=DEV), -1)[0]
    P = N * S
    fl_f = 2.0 * {8: 593408, 4: 315136}[D] * P
    fl_d = 2.0 * {8: 557696, 4: 295552}[D] * P
    t_inf = timeit(lambda: net._run_forward("rays", rb, z, P, keep=False))
    t_trn = timeit(lambda: net._run_forward("rays", rb, z, P, keep=True))
    out, saved = net._run_forward("rays", rb, z, P, keep=True)
    d_out = torch.randn(P, 4, device=DEV)
    st = net._state(); plan = net._plan
    n_tiles = (P + 127) // 128
    stash_b = torch.empty(n_tiles * plan.bwd_slots * L.SLAB_BYTES, device=DEV, dtype=torch.uint8)
    def dgrad(keep=True):
        args = L.ChainArgs(); args.P = P
        args.wblob, args.fblob = st["wb"].data_ptr(), st["flat"].data_ptr()
        args.d_out, args.masks = d_out.data_ptr(), saved[1].data_ptr()
        args.stash = stash_b.data_ptr() if keep else None
        L.check(L.lib().dln_mlp_chain(C.byref(plan.bwd), C.byref(args), st["sms"], dn.ops._stream()), "dgrad")
    t_dg = timeit(dgrad)
    t_dg_nostash = timeit(lambda: dgrad(False))
    gflat = torch.zeros(plan.n_params, device=DEV)
    res = {}
    _part = torch.empty(len(plan.wgrad) * 64 * L.WGRAD_PARTIAL_FLOATS, device=DEV) if not os.environ.get("DLN_WGRAD_ATOMICS") else None
    PARTIAL = _part.data_ptr() if _part is not None else None
    for splits in [int(x) for x in os.environ.get("SPLITS", "10,21,32,42").split(",")]:
        res[splits] = timeit(lambda: L.check(L.lib().dln_mlp_wgrad(st["items"].data_ptr(), len(plan.wgrad), splits, saved[0].data_ptr(), plan.fwd_slots, stash_b.data_ptr(), plan.bwd_slots, n_tiles, gflat.data_ptr(), PARTIAL, dn.ops._stream()), "wgrad"))
    print("D=%d P=%d  fwd(infer) %.3f ms %.0f TF/s | fwd(train) %.3f ms %.0f TF/s | dgrad %.3f ms %.0f TF/s (no stash %.3f) | wgrad %s" % (
        D, P, t_inf, fl_f / t_inf / 1e9, t_trn, fl_f / t_trn / 1e9, t_dg, fl_d / t_dg / 1e9, t_dg_nostash,
        {k: "%.3f" % v for k, v in res.items()}))
