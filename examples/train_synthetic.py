"""Minimal training loop in the shape of the reference's train() (run_nerf.py:1328-1420, :1500-1536, :1759-1774,
:1843-1847) on the B200 path, with a synthetic scene instead of a dataset (none ships with the reference):

    colour target  = 0.5 + 0.5 * unit view direction      (a view-dependent "sky")
    depth  target  = 0.7 in NDC for the depth rays         (a fronto-parallel wall)
    class  target  = quadrant of the view direction        (--semantic: 4 classes, cross-entropy on sem_preds / sem_preds0
                                                            with semantic_lambda = 0.01, fern_dsnerf.txt:55-56)

It uses the optional swaps of INTEGRATION.md: DeviceRayLoader for RayDataset + DataLoader, FlatAdam for
torch.optim.Adam, GraphedTrainStep for render + loss + backward.  Prints loss / PSNR every 50 iterations.

    python examples/train_synthetic.py [--iters 300] [--n-rand 1024] [--drop-in] [--semantic]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlnerf_b200 as dn  # noqa: E402

H, W, FOCAL = 94, 352, 138.14          # KITTI-360 at factor 4 (configs/fern_dsnerf.txt:22, :50-51)


def synthetic_rays(n_views=8, seed=0):
    """[N, ro+rd+rgb, 3] like the reference's rays_rgb array (run_nerf.py:1126-1140), from jittered forward poses."""
    rs = np.random.RandomState(seed)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - W * .5) / FOCAL, -(j - H * .5) / FOCAL, -np.ones_like(i)], -1)
    out = []
    for _ in range(n_views):
        t = np.array([rs.uniform(-.3, .3), rs.uniform(-.05, .05), rs.uniform(-.3, .3)], np.float32)
        rd = dirs.reshape(-1, 3)
        ro = np.broadcast_to(t, rd.shape)
        unit = rd / np.linalg.norm(rd, axis=-1, keepdims=True)
        out.append(np.stack([ro, rd, 0.5 + 0.5 * unit], 1))
    return np.concatenate(out, 0).astype(np.float32)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--n-rand", type=int, default=1024)
    ap.add_argument("--lrate", type=float, default=5e-4)
    ap.add_argument("--drop-in", action="store_true", help="render() + img2mse + loss.backward() instead of the graphed step")
    ap.add_argument("--semantic", action="store_true", help="networks with a 4-class semantic head + cross-entropy loss")
    args = ap.parse_args(argv)
    K, slam = (4, 0.01) if args.semantic else (None, 0.)
    dev = torch.device("cuda")
    torch.manual_seed(3407)
    model = dn.NeRF(D=4, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True, semantic_num_classes=K).to(dev)
    model_fine = dn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True, semantic_num_classes=K).to(dev)
    optimizer = dn.FlatAdam([model, model_fine], lr=args.lrate, betas=(0.9, 0.999))
    n_rgb = args.n_rand // 2
    n_dep = args.n_rand - n_rgb
    rays = synthetic_rays()
    rgb_loader = dn.DeviceRayLoader(rays, batch_size=n_rgb, device=dev)
    dep_loader = dn.DeviceRayLoader(rays, batch_size=n_dep, device=dev)
    rgb_it, dep_it = iter(rgb_loader), iter(dep_loader)
    target_depth = torch.full((n_dep,), 0.7, device=dev)
    kw = dict(N_samples=64, N_importance=64, perturb=1., raw_noise_std=1., depth_lambda=0.1, depth_importance=1.)
    if K:
        kw["semantic_lambda"] = slam
    step = None if args.drop_in else dn.GraphedTrainStep(H, W, FOCAL, args.n_rand, n_rgb, model, model_fine, **kw)
    q = dn.FusedQuery(dn.get_embedder(10, 0)[0], dn.get_embedder(4, 0)[0], 65536, 10, 4, 0)
    losses = []
    for i in range(args.iters):
        def nxt(it, loader):
            try:
                b = next(it)
                if b.shape[0] == loader.batch_size:
                    return b, it
            except StopIteration:
                pass
            it = iter(loader)               # the reference's restart idiom (run_nerf.py:1335-1340)
            return next(it), it
        b_rgb, rgb_it = nxt(rgb_it, rgb_loader)
        b_dep, dep_it = nxt(dep_it, dep_loader)
        batch = torch.cat([b_rgb, b_dep], 0).transpose(0, 1)         # [3, N_rand, 3]
        batch_rays, target_s = batch[:2].contiguous(), batch[2, :n_rgb].contiguous()
        target_sem = None
        if K:                                   # the loader of run_nerf.py:1202 yields (rays, class index) pairs
            rd = batch_rays[1, :n_rgb]
            target_sem = (rd[:, 0] > 0).long() + 2 * (rd[:, 1] > 0).long()
        sem_ce = None
        if step is not None:
            out = step(batch_rays, target_s, target_depth, target_semantic=target_sem)
            loss, psnr = out["loss"], out["psnr"]
            sem_ce = out.get("semantic_loss")
        else:
            rgb, disp, acc, depth, extras = dn.render(H, W, FOCAL, chunk=1 << 20, rays=batch_rays, retraw=True,
                                                      network_query_fn=q, perturb=1., N_importance=64,
                                                      network_fine=model_fine, N_samples=64, network_fn=model,
                                                      use_viewdirs=True, white_bkgd=False, raw_noise_std=1., ndc=True,
                                                      near=0., far=1., semantic_loss=bool(K))
            optimizer.zero_grad()
            img_loss = dn.img2mse(rgb[:n_rgb], target_s)
            loss = img_loss + 0.1 * dn.img2mse(depth[n_rgb:], target_depth) + dn.img2mse(extras['rgb0'][:n_rgb], target_s)
            if K:                               # run_nerf.py:1541-1548
                sem_ce = torch.nn.functional.cross_entropy(extras['sem_preds'][:n_rgb], target_sem)
                loss = loss + slam * (sem_ce + torch.nn.functional.cross_entropy(extras['sem_preds0'][:n_rgb], target_sem))
            loss.backward()
            psnr = dn.mse2psnr(img_loss)
        optimizer.step()
        new_lrate = args.lrate * (0.1 ** (i / 250000.))                 # run_nerf.py:1843-1847
        for g in optimizer.param_groups:
            g['lr'] = new_lrate
        if i % 50 == 0 or i == args.iters - 1:
            losses.append((float(loss.detach()), float(psnr.detach())) + ((float(sem_ce.detach()),) if K else ()))
            print("iter %4d  loss %.5f  psnr %.2f%s" % (i, losses[-1][0], losses[-1][1],
                                                       "  semantic CE %.4f" % losses[-1][2] if K else ""), flush=True)
    return losses


if __name__ == "__main__":
    main()
