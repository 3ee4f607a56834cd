"""Does a read-bound kernel next to a write-bound one beat running them back to back?  Fine-net forward (stash writes,
2.6 GB) on `a` SMs on one stream, fine-net wgrad (stash reads, 4.95 GB) of an independent batch on the other SMs on a
second stream; CUDA events around both, against the two launched back to back on all SMs.  524 288 points each."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
L = dn._lib
DEV = "cuda"
N, S, D = 4096, 128, 8
torch.manual_seed(0)
net = dn.NeRF(D=D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(DEV)
rb = torch.randn(N, 11, device=DEV); rb[:, 6] = 0; rb[:, 7] = 1
z = torch.sort(torch.rand(N, S, device=DEV), -1)[0]
P = N * S
st = net._state(); plan = net._plan
n_tiles = P // 128
out, saved = net._run_forward("rays", rb, z, P, keep=True)          # batch 1: operands of the wgrad
d_out = torch.randn(P, 4, device=DEV)
stash_b = torch.empty(n_tiles * plan.bwd_slots * L.SLAB_BYTES, device=DEV, dtype=torch.uint8)
args = L.ChainArgs(); args.P = P
args.wblob, args.fblob = st["wb"].data_ptr(), st["flat"].data_ptr()
args.d_out, args.masks, args.stash = d_out.data_ptr(), saved[1].data_ptr(), stash_b.data_ptr()
L.check(L.lib().dln_mlp_chain(C.byref(plan.bwd), C.byref(args), st["sms"], dn.ops._stream()), "dgrad")
gflat = torch.zeros(plan.n_flat, device=DEV)
n_items = len(plan.wgrad)
part = torch.empty(n_items * 16 * L.WGRAD_PARTIAL_FLOATS, device=DEV)
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def fwd(sms):
    net._run_forward("rays", rb, z, P, keep=True, sms=sms)          # batch 2 (same inputs, fresh stash)


def wgrad(sms):
    splits = max(1, min(n_tiles, sms // n_items))
    L.check(L.lib().dln_mlp_wgrad(st["items"].data_ptr(), n_items, splits, saved[0].data_ptr(), plan.fwd_slots,
                                  stash_b.data_ptr(), plan.bwd_slots, n_tiles, gflat.data_ptr(), part.data_ptr(),
                                  dn.ops._stream()), "wgrad")


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def serial():
    fwd(148)
    wgrad(148)


def both(a):
    def run():
        main = torch.cuda.current_stream()
        s1.wait_stream(main); s2.wait_stream(main)
        with torch.cuda.stream(s2):
            wgrad(148 - a)
        with torch.cuda.stream(s1):
            fwd(a)
        main.wait_stream(s1); main.wait_stream(s2)
    return run


print("fwd alone %.3f ms, wgrad alone %.3f ms, back to back %.3f ms" % (timed(lambda: fwd(148)), timed(lambda: wgrad(148)), timed(serial)))
for a in (64, 80, 96, 108, 120):
    print("fwd on %3d SMs | wgrad on %3d SMs (%2d splits): %.3f ms   (fwd alone on %d: %.3f, wgrad alone on %d: %.3f)" % (
        a, 148 - a, (148 - a) // n_items, timed(both(a)), a, timed(lambda: fwd(a)), 148 - a, timed(lambda: wgrad(148 - a))))
