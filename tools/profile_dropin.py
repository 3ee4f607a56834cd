"""Host/device breakdown of one drop-in training step (render + img2mse + loss.backward) with torch.profiler."""
import os, sys, time
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
rays, tgt, dep = torch.stack([ro, rd], 0).to(dev), tgt.to(dev), dep.to(dev)

def step():
    rgb, disp, acc, depth, extras = dn.render(bench.H, bench.W, bench.FOCAL, chunk=1 << 30, rays=rays, retraw=True, **kw)
    for p in params:
        p.grad = None
    loss = dn.img2mse(rgb[:n_rgb], tgt) + 0.01 * dn.img2mse(depth[n_rgb:], dep) + dn.img2mse(extras["rgb0"][:n_rgb], tgt)
    loss.backward()
    return float(loss.item())

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print("wall per step %.3f ms" % ((time.perf_counter() - t0) * 100))
# where the host time goes, without the profiler: time until the last launch is queued
t0 = time.perf_counter()
rgb, disp, acc, depth, extras = dn.render(bench.H, bench.W, bench.FOCAL, chunk=1 << 30, rays=rays, retraw=True, **kw)
t1 = time.perf_counter()
loss = dn.img2mse(rgb[:n_rgb], tgt) + 0.01 * dn.img2mse(depth[n_rgb:], dep) + dn.img2mse(extras["rgb0"][:n_rgb], tgt)
t2 = time.perf_counter()
loss.backward()
t3 = time.perf_counter()
torch.cuda.synchronize()
t4 = time.perf_counter()
print("host: render() returns after %.3f ms, loss built +%.3f, backward() returns +%.3f, GPU drained +%.3f" % (
    (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=50))
