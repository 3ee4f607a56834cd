"""GPU diagnostic: many tiles per persistent CTA.  Compares every backward-stash slot per tile and the
final gradients against a bf16-emulating torch reference computed on the GPU."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import dlnerf_b200 as dn
from gpu_util import O, _ste, bf16r, make_net, read_stash, stash_rows, cosine, rel_l2
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
D = 8
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 400
net, params, spec = make_net(D)
g = torch.Generator().manual_seed(0)
pts = torch.rand(P, 3, generator=g) * 2 - 1
dirs = torch.randn(P, 3, generator=g); dirs = dirs / dirs.norm(dim=-1, keepdim=True)
x = torch.cat([O.posenc(pts, 10), O.posenc(dirs, 4)], -1)
cot = torch.randn(P, 4, generator=g)
# reference on GPU
pl = {k: v.clone().to(DEV).requires_grad_(True) for k, v in params.items()}
xg = x.to(DEV)
xp, xd = _ste(xg[:, :63]), _ste(xg[:, 63:])
zs, hs, h = [], [], xp
for i in range(D):
    z = h @ _ste(pl["pts_linears.%d.weight" % i]).T + pl["pts_linears.%d.bias" % i]; z.retain_grad(); zs.append(z)
    h32 = torch.relu(z); h = _ste(h32); hs.append(h)
    if i in spec.skips: h = torch.cat([xp, h], -1)
sigma = h32 @ pl["alpha_linear.weight"].T + pl["alpha_linear.bias"]
feat = h @ _ste(pl["feature_linear.weight"]).T + pl["feature_linear.bias"]; feat.retain_grad()
zv = torch.cat([_ste(feat), xd], -1) @ _ste(pl["views_linears.0.weight"]).T + pl["views_linears.0.bias"]; zv.retain_grad()
rgb = torch.relu(zv) @ pl["rgb_linear.weight"].T + pl["rgb_linear.bias"]
(torch.cat([rgb, sigma], -1) * cot.to(DEV)).sum().backward()

L = dn._lib
out, saved = net._run_forward("x", xg, None, P, keep=True)
st = net._state(); plan = net._plan
n_tiles = (P + 127) // 128
sf = read_stash(saved[0], n_tiles, plan.fwd_slots)
def per_tile(name, got, ref):
    got = got.float().cpu(); ref = ref.detach().float().cpu()
    T = got.shape[0] // 128
    e = (got[:T*128] - ref[:T*128]).reshape(T, -1).norm(dim=1) / (ref[:T*128].reshape(T, -1).norm(dim=1) + 1e-30)
    bad = (e > 0.05).nonzero().flatten().tolist()
    print("%-14s overall relL2 %.3e  bad tiles %d/%d  first bad %s" % (name, rel_l2(got, ref), len(bad), T, bad[:12]))
for i in range(D):
    per_tile("fwd H%d" % i, stash_rows(sf, 2 + 4 * i, 4, P), hs[i])
per_tile("fwd feat", stash_rows(sf, 2 + 4 * D, 4, P), bf16r(feat))
stash_b = torch.zeros(n_tiles * plan.bwd_slots * L.SLAB_BYTES, device=DEV, dtype=torch.uint8)
args = L.ChainArgs(); args.P = P
d = cot.to(DEV).contiguous()
args.wblob, args.fblob = st["wb"].data_ptr(), st["flat"].data_ptr()
args.d_out, args.stash, args.masks = d.data_ptr(), stash_b.data_ptr(), saved[1].data_ptr()
L.check(L.lib().dln_mlp_chain(C.byref(plan.bwd), C.byref(args), st["sms"], dn.ops._stream()), "dgrad")
torch.cuda.synchronize()
sb = read_stash(stash_b, n_tiles, plan.bwd_slots)
per_tile("dZ views", stash_rows(sb, 1, 2, P), zv.grad)
per_tile("d feature", stash_rows(sb, 3, 4, P), feat.grad)
for l in range(D - 1, -1, -1):
    per_tile("dZ layer %d" % l, stash_rows(sb, 7 + 4 * (D - 1 - l), 4, P), zs[l].grad)
# wgrad
gflat = torch.zeros(plan.n_params, device=DEV)
n_items = len(plan.wgrad)
for splits in (max(1, min(n_tiles, (2 * st["sms"]) // n_items)), n_tiles):
    gflat.zero_()
    PARTIAL = None
    L.check(L.lib().dln_mlp_wgrad(st["items"].data_ptr(), n_items, splits, saved[0].data_ptr(), plan.fwd_slots,
                                  stash_b.data_ptr(), plan.bwd_slots, n_tiles, gflat.data_ptr(), PARTIAL, dn.ops._stream()), "wgrad")
    torch.cuda.synchronize()
    print("wgrad with splits=%d (tiles per CTA %.1f)" % (splits, n_tiles / splits))
    for name, shp in plan.shape.param_shapes():
        o = plan.offsets[name]; n = pl[name].numel()
        got = gflat[o:o + n].view(pl[name].shape)
        print("   %-26s cos %.5f relL2 %.3e" % (name, cosine(got, pl[name].grad), rel_l2(got, pl[name].grad)))
