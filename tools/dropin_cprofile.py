"""Host-side profile (cProfile) of the drop-in step: where the Python time between kernel launches goes."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
import bench

dev = torch.device("cuda:0")
torch.manual_seed(3407)
net_c = dn.NeRF(D=4, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(dev)
net_f = dn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(dev)
params = list(net_c.parameters()) + list(net_f.parameters())
q = dn.FusedQuery(dn.get_embedder(10, 0)[0], dn.get_embedder(4, 0)[0], 65536, 10, 4, 0)
kw = dict(network_query_fn=q, perturb=1.0, N_importance=64, network_fine=net_f, N_samples=64, network_fn=net_c,
          use_viewdirs=True, white_bkgd=False, raw_noise_std=1.0, ndc=True, near=0., far=1.)
ro, rd, tgt, dep, n_rgb, n_dep = bench.make_batch(4096, 3407)
h_rays, h_tgt, h_dep = torch.stack([ro, rd], 0).pin_memory(), tgt.pin_memory(), dep.pin_memory()

def step():
    rays, t, d = h_rays.to(dev, non_blocking=True), h_tgt.to(dev, non_blocking=True), h_dep.to(dev, non_blocking=True)
    rgb, disp, acc, depth, extras = dn.render(bench.H, bench.W, bench.FOCAL, chunk=1 << 30, rays=rays, retraw=True, **kw)
    for p in params:
        p.grad = None
    loss = dn.img2mse(rgb[:n_rgb], t) + 0.01 * dn.img2mse(depth[n_rgb:], d) + dn.img2mse(extras["rgb0"][:n_rgb], t)
    loss.backward()
    return float(loss.item())

for _ in range(10):
    step()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
