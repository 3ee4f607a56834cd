"""What an fp8 activation / dZ stash would do to the weight gradients (CPU study, no kernels): the fine network's
per-layer wgrad dW_l = dZ_l^T H_{l-1} with the operands rounded to bf16 (what the kernels stash today), to e4m3 / e5m2
(per-tensor power-of-two scale for dZ), against fp32 operands.  Reported: rel-L2 of every dW_l and the aggregate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nerf_oracle as O

torch.manual_seed(0)
N, S = int(os.environ.get("RAYS", 384)), 128
spec = O.MLPSpec(D=8)
p = O.trained_like(O.init_params(spec, 3407 + 8), 1.0)
ro, rd = O.synth_rays(N, seed=3)
rb = O.pack_rays(378, 504, 407.6, ro, rd)
z = torch.sort(torch.rand(N, S), -1)[0]
pts = rb[:, None, 0:3] + rb[:, None, 3:6] * z[..., None]
x = torch.cat([O.posenc(pts.reshape(-1, 3), 10), O.posenc(rb[:, None, -3:].expand(-1, S, -1).reshape(-1, 3), 4)], -1)

# manual forward keeping the layer inputs H and pre-activations Z of the 8 trunk layers
W = [p["pts_linears.%d.weight" % i] for i in range(8)]
b = [p["pts_linears.%d.bias" % i] for i in range(8)]
H, Z = [], []
h = x[:, :63]
for i in range(8):
    inp = h if i != 5 else torch.cat([x[:, :63], h], -1)
    H.append(inp)
    zz = (inp @ W[i].T + b[i]).requires_grad_(True)
    Z.append(zz)
    h = torch.relu(zz)
sigma = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
feat = h @ p["feature_linear.weight"].T + p["feature_linear.bias"]
hv = torch.relu(torch.cat([feat, x[:, 63:]], -1) @ p["views_linears.0.weight"].T + p["views_linears.0.bias"])
rgb = hv @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
raw = torch.cat([rgb, sigma], -1).reshape(N, S, 4)
out = O.raw2outputs(raw, z, rb[:, 3:6], None, False)
tgt = torch.rand(N, 3)
loss = torch.mean((out[0] - tgt) ** 2) + 0.01 * torch.mean((out[4] - torch.rand(N)) ** 2)
dZ = torch.autograd.grad(loss, Z)

def q_bf16(t): return t.to(torch.bfloat16).float()
def q_fp8(t, dt, scaled):
    if not scaled:
        return t.to(dt).float()
    amax = t.abs().max().clamp_min(1e-30)
    top = 448.0 if dt == torch.float8_e4m3fn else 57344.0
    s = 2.0 ** torch.floor(torch.log2(top / amax))
    return (t * s).to(dt).float() / s

schemes = {
    "bf16 x bf16 (today)": (q_bf16, q_bf16),
    "e4m3 H, e4m3 dZ (scaled)": (lambda t: q_fp8(t, torch.float8_e4m3fn, False), lambda t: q_fp8(t, torch.float8_e4m3fn, True)),
    "e4m3 H, e5m2 dZ (scaled)": (lambda t: q_fp8(t, torch.float8_e4m3fn, False), lambda t: q_fp8(t, torch.float8_e5m2, True)),
    "e4m3 H, bf16 dZ": (lambda t: q_fp8(t, torch.float8_e4m3fn, False), q_bf16),
}
exact = [dZ[i].T.double() @ H[i].double() for i in range(8)]
print("%d points; |dW_l| fp32: %s" % (x.shape[0], " ".join("%.2e" % e.norm() for e in exact)))
for name, (qh, qz) in schemes.items():
    errs, num, den = [], 0.0, 0.0
    for i in range(8):
        got = qz(dZ[i]).T.double() @ qh(H[i]).double()
        errs.append(float((got - exact[i]).norm() / exact[i].norm()))
        num += float((got - exact[i]).pow(2).sum()); den += float(exact[i].pow(2).sum())
    print("%-28s aggregate rel-L2 %.2e | per layer %s" % (name, (num / den) ** 0.5, " ".join("%.1e" % e for e in errs)))
