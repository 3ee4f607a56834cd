"""Print the in-kernel timeline of CTA 0 for the forward chain (debug)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
L = dn._lib
DEV = "cuda"
N, S, D = 4096, 128, 8
keep = int(os.environ.get("KEEP", "1"))
net = dn.NeRF(D=D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(DEV)
rb = torch.randn(N, 11, device=DEV); rb[:, 6] = 0; rb[:, 7] = 1
z = torch.sort(torch.rand(N, S, device=DEV), -1)[0]
P = N * S
st = net._state(); net._pack(st); pl = net._plan
n_tiles = P // 128
out = torch.empty(P, 4, device=DEV)
stash = torch.empty(n_tiles * pl.fwd_slots * L.SLAB_BYTES, device=DEV, dtype=torch.uint8)
masks = torch.empty(pl.mask_slots * n_tiles * 1024, device=DEV, dtype=torch.int32)
trace = torch.zeros(4 * 64 * 8, device=DEV, dtype=torch.int64)
args = L.ChainArgs(); args.P = P
args.rays, args.ray_stride, args.vd_col = rb.data_ptr(), 11, 8
args.z, args.S = z.data_ptr(), S
args.wblob, args.fblob, args.out = st["wf"].data_ptr(), st["flat"].data_ptr(), out.data_ptr()
if keep:
    args.stash, args.masks = stash.data_ptr(), masks.data_ptr()
for it in range(3):
    args.trace = trace.data_ptr() if it == 2 else None
    L.check(L.lib().dln_mlp_chain(C.byref(pl.fwd), C.byref(args), st["sms"], dn.ops._stream()), "fwd")
torch.cuda.synchronize()
t = trace.cpu()[:64].view(4, 2, 8) & 0xFFFFFFFF
t0 = int(t[0, 0, 0])
rel = lambda x: (int(x) - t0) & 0xFFFFFFFF if int(x) else -1
print("keep=%d; gsteps 12,13 of CTA 0 (SM clocks, smem trace)" % keep)
print("MMA  [start, a_ready j0, j1, j2, j3, j4, commit0, end]")
print("EPI  [before wait, acc_full seen, ld0 done, computed c0, ld_done passed, s_free passed, arrived c0, arrived c1]")
for gs in range(2):
    print("g%02d MMA %s\n    WG0 %s\n    WG3 %s\n    STASH[a_ready seen slab0..3 | s_free arrived slab0..3] %s" % (12 + gs, [rel(x) for x in t[0, gs]], [rel(x) for x in t[1, gs]], [rel(x) for x in t[2, gs]], [rel(x) for x in t[3, gs]]))
