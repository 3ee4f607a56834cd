"""Small forward + backward through the CTA-pair chain kernels for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
torch.manual_seed(0)
for D, vd, W in ((8, True, 256), (4, False, 256), (8, True, 128)):
    net = dn.NeRF(D=D, W=W, input_ch=63, input_ch_views=27, use_viewdirs=vd).cuda()
    N, S = 37, 24                                  # 888 points: 7 tiles, the last one partial, two pairs
    rb = torch.randn(N, 11, device="cuda"); rb[:, 6] = 0; rb[:, 7] = 1
    z = torch.sort(torch.rand(N, S, device="cuda"), -1)[0]
    raw = net.forward_rays(rb, z)
    raw.sum().backward()
    zg = torch.empty(N, S, device="cuda")
    raw2 = net.forward_rays(rb, zg, strat=dict(rng=None, lindisp=False))
    torch.cuda.synchronize()
    print("D=%d vd=%s W=%d ok" % (D, vd, W), float(raw.abs().sum()), float(net.pts_linears[0].weight.grad.abs().sum()))
