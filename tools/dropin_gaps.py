"""GPU idle gaps of the drop-in step (render + loss + backward from pinned host buffers, loss read back): kernel timeline
from torch.profiler, gaps > 15 us between consecutive kernels (any stream) printed with their neighbours."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlnerf_b200 as dn
import bench
from torch.profiler import profile, ProfilerActivity

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

for _ in range(8):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
end = 0
busy = 0
last = None
print("kernels: %d over %.3f ms" % (len(evs), (evs[-1].time_range.end - t0) / 1e3))
for e in evs:
    s, t = e.time_range.start - t0, e.time_range.end - t0
    if last is not None and s - end > 15:
        print("  idle %6.0f us  after %-40s before %-40s at %.3f ms" % (s - end, last.name[:40], e.name[:40], s / 1e3))
    if t > end:
        busy += t - max(s, end)
        end, last = t, e
print("busy %.3f ms of %.3f ms (4 steps)" % (busy / 1e3, end / 1e3))
