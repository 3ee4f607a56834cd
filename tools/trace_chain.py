"""Print the in-kernel timeline of CTA 0 (debug): 16 consecutive layer steps of the forward (default) or dgrad
(BWD=1) chain, whole tiles with their boundaries.  Columns are SM clocks relative to the first traced event."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
L = dn._lib
DEV = "cuda"
N, S, D = 4096, int(os.environ.get("S", 128)), int(os.environ.get("D", 8))
keep = int(os.environ.get("KEEP", "1"))
bwd = int(os.environ.get("BWD", "0"))
net = dn.NeRF(D=D, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(DEV)
rb = torch.randn(N, 11, device=DEV); rb[:, 6] = 0; rb[:, 7] = 1
z = torch.sort(torch.rand(N, S, device=DEV), -1)[0]
P = N * S
st = net._state(); net._pack(st); pl = net._plan
n_tiles = P // 128
out = torch.empty(P, 4, device=DEV)
stash = torch.empty(n_tiles * pl.fwd_slots * L.SLAB_BYTES, device=DEV, dtype=torch.uint8)
stash_b = torch.empty(n_tiles * pl.bwd_slots * L.SLAB_BYTES, device=DEV, dtype=torch.uint8)
masks = torch.zeros(pl.mask_slots * n_tiles * 1024, device=DEV, dtype=torch.int32)
trace = torch.zeros(256, device=DEV, dtype=torch.int64)
args = L.ChainArgs(); args.P = P
if not bwd:
    args.rays, args.ray_stride, args.vd_col = rb.data_ptr(), 11, 8
    args.z, args.S = z.data_ptr(), S
    args.wblob, args.fblob, args.out = st["wf"].data_ptr(), st["flat"].data_ptr(), out.data_ptr()
    if keep:
        args.stash, args.masks = stash.data_ptr(), masks.data_ptr()
    prog = pl.fwd
else:
    d_out = torch.randn(P, 4, device=DEV)
    args.wblob, args.fblob = st["wb"].data_ptr(), st["flat"].data_ptr()
    args.d_out, args.masks = d_out.data_ptr(), masks.data_ptr()
    if keep:
        args.stash = stash_b.data_ptr()
    prog = pl.bwd
for it in range(3):
    args.trace = trace.data_ptr() if it == 2 else None
    L.check(L.lib().dln_mlp_chain(C.byref(prog), C.byref(args), st["sms"], dn.ops._stream()), "chain")
torch.cuda.synchronize()
n_steps = prog.n_steps
if os.environ.get("DLN_CHAIN", "2") != "1":
    # CTA-pair kernel: 8 consecutive (round, step) pairs from index 10 on, per slot: MMA first issue / commit issued,
    # epilogue thread 0: accumulator seen / stores done / hand-off (+ next prologue) done
    t = (trace.cpu()[:128] & 0xFFFF).view(8, 16)
    if not any(int(x) for x in t[0]):
        print("no timeline recorded (build with DLN_NVCC_EXTRA=-DDLN_CHAIN2_TRACE for it)")
        sys.exit(0)
    t0 = min(int(x) for x in t[0] if int(x))
    rel = lambda x: (((int(x) - t0) & 0xFFFF) << 3) if int(x) else -1
    print("%s D=%d keep=%d: %d steps per tile; (round, step) index 10..17 of CTA 0 (leader of pair 0)" % ("dgrad" if bwd else "fwd", D, keep, n_steps))
    print("  idx s | X: mma-first mma-commit  acc-seen stores-done handed-off | Y: mma-first mma-commit  acc-seen stores-done handed-off")
    for i in range(8):
        r = [rel(x) for x in t[i]]
        print("%5d %d | %10d %10d %9d %11d %10d | %10d %10d %9d %11d %10d" % (
            10 + i, (10 + i) % n_steps, r[0], r[1], r[2], r[3], r[4], r[8], r[9], r[10], r[11], r[12]))
    print("epilogue thread 0, cycles after acc-seen: first tmem load landed / chunk 0 converted / second load landed (chunk 0 stored) / stores done / handed off")
    for i in range(8):
        r = [rel(x) for x in t[i]]
        f = lambda o: " ".join("%6d" % (r[o + k] - r[o + 2]) for k in (5, 6, 7, 3, 4))
        print("%5d %d | X: %s | Y: %s" % (10 + i, (10 + i) % n_steps, f(0), f(8)))
    sys.exit(0)
t = (trace.cpu()[:128] & 0xFFFF).view(16, 8)
t0 = int(t[0, 0])
rel = lambda x: (((int(x) - t0) & 0xFFFF) << 3) if int(x) else -1
print("%s D=%d keep=%d: %d steps per tile; gsteps 8..23 of CTA 0" % ("dgrad" if bwd else "fwd", D, keep, n_steps))
print("gstep s | MMA: enter  ops-ready  issued | EPI: acc-seen  c0-handoff  c1-handoff  staged  step-left | acc period")
prev = None
for i in range(16):
    g = 8 + i
    r = [rel(x) for x in t[i]]
    per = (r[3] - prev) if prev is not None and r[3] >= 0 else 0
    prev = r[3] if r[3] >= 0 else prev
    print("%5d %d | %8d %8d %8d | %8d %8d %8d %8d %8d | %6d" % (g, g % n_steps, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], per))
