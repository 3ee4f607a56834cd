import os, sys, torch
sys.path.insert(0, os.getcwd())
import dlnerf_b200 as dn, bench
dev = torch.device("cuda:0")
torch.manual_seed(3407)
net_c = dn.NeRF(D=4, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(dev)
net_f = dn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True).to(dev)
ro, rd, tgt, dep, n_rgb, n_dep = bench.make_batch(4096, 3407)
rays, tgt, dep = torch.stack([ro, rd], 0).to(dev), tgt.to(dev), dep.to(dev)
def run(**kw):
    g = dn.GraphedTrainStep(bench.H, bench.W, bench.FOCAL, 4096, n_rgb, net_c, net_f, N_samples=64, N_importance=64,
                            perturb=1., raw_noise_std=1., depth_lambda=0.01, depth_importance=1., **kw)
    for _ in range(5): g(rays, tgt, dep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): g(rays, tgt, dep)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30
print("no overlap        %.3f ms" % run(overlap_coarse_backward=False))
for k in (None, 24, 32, 48, 64, 96):
    print("overlap, coarse dgrad on %s SMs: %.3f ms" % (k, run(coarse_sms=k)))
