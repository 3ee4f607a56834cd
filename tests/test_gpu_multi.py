"""Hardware check of the multi-GPU maths (needs >= 2 GPUs; skipped on a single-GPU box -- run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):

* NCCL-summed gradients of two ray shards (`train_step(world_size=2)`, eager and as a captured graph with the
  all-reduces inside it) == the single-GPU gradients of the union batch, per tensor rel-L2 <= 2e-3 (identical kernels;
  fp32 atomics ordering and the order of the sum over the ranks differ),
* `render_patch_nograd_sharded` over NCCL == the single-GPU render of the whole patch, bit for bit.

Randomness is injected (`_rng`) or off so both sides see the same draws.  SURVEY.md section 8(e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpu_util import O, dn, make_net, rel_l2

pytestmark = pytest.mark.gpu
H, W, FOCAL = 378, 504, 407.6
N_RGB, N_DEP = 192, 128                 # per-rank shares 96 + 64 rays
KW = dict(N_samples=64, N_importance=64, depth_lambda=0.01, depth_importance=0.7)


def _batch(seed=51):
    ro, rd = O.synth_rays(N_RGB + N_DEP, seed=seed)
    rng = O.synth_rng(N_RGB + N_DEP, 64, 64, seed=seed)
    tgt, dep = O.synth_targets(N_RGB, N_DEP, seed=seed)
    return torch.stack([ro, rd], 0), tgt, dep, {k: getattr(rng, k) for k in ("t_rand", "noise0", "u", "noise1")}


def _nets(dev):
    net_c, _, _ = make_net(4, seed=61, device=dev)
    net_f, _, _ = make_net(8, seed=62, device=dev)
    return net_c, net_f


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    d = dn()
    rays, tgt, dep, inj = _batch()
    net_c, net_f = _nets(dev)
    r, t, dp, _, n_loc = d.shard_ray_batch(rays, tgt, dep, N_RGB, rank, world)
    a0, a1 = d.shard_bounds(N_RGB, rank, world)
    b0, b1 = d.shard_bounds(N_DEP, rank, world)
    inj_loc = {k: torch.cat([v[a0:a1], v[N_RGB + b0:N_RGB + b1]], 0).to(dev) for k, v in inj.items()}
    res = {}
    o = d.train_step(H, W, FOCAL, r.to(dev), t.to(dev), dp.to(dev), n_loc, net_c, net_f, perturb=1., raw_noise_std=1.,
                     world_size=world, global_counts=(N_RGB, N_DEP), _rng=inj_loc, **KW)
    res["eager"] = [p.grad.detach().cpu().clone() for p in list(net_c.parameters()) + list(net_f.parameters())]
    res["eager_loss_local"] = float(o["loss"])
    # captured graph with the two all-reduces inside (no injected draws in a graph: randomness off)
    step = d.GraphedTrainStep(H, W, FOCAL, r.shape[1], n_loc, net_c, net_f, world_size=world,
                              global_counts=(N_RGB, N_DEP), perturb=0., raw_noise_std=0., **KW)
    step(r.to(dev), t.to(dev), dp.to(dev))
    step(r.to(dev), t.to(dev), dp.to(dev))
    res["graph"] = [p.grad.detach().cpu().clone() for p in list(net_c.parameters()) + list(net_f.parameters())]
    step.close()                  # the graph holds NCCL kernels: release it before the process group goes away
    # the no-grad part of a patch render, split over the ranks and all-gathered
    q = d.FusedQuery(d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0], 1 << 16, 10, 4, 0)
    rk = dict(network_query_fn=q, perturb=0., N_importance=64, network_fine=net_f, N_samples=64, network_fn=net_c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0., ndc=True, near=0., far=1.)
    keys = ["rgb_map", "depth_map", "rgb0"]
    got = d.render_patch_nograd_sharded(H, W, FOCAL, (rays[0, :301].to(dev), rays[1, :301].to(dev)), rank, world,
                                        keep_keys=keys, **rk)
    res["patch"] = {k: v.cpu() for k, v in got.items()}
    torch.cuda.synchronize()
    if rank == 1:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_sharded_step_equals_the_single_gpu_step(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "multi.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)

    d = dn()
    dev = "cuda:0"
    rays, tgt, dep, inj = _batch()
    net_c, net_f = _nets(dev)
    nets = list(net_c.parameters()) + list(net_f.parameters())
    d.train_step(H, W, FOCAL, rays.to(dev), tgt.to(dev), dep.to(dev), N_RGB, net_c, net_f, perturb=1., raw_noise_std=1.,
                 _rng={k: v.to(dev) for k, v in inj.items()}, **KW)
    worst = max(rel_l2(g, p.grad) for g, p in zip(got["eager"], nets))
    print("  NCCL-summed shard gradients vs the union batch on one GPU (live draws injected): worst rel-L2 %.3e" % worst)
    assert worst <= 2e-3
    d.train_step(H, W, FOCAL, rays.to(dev), tgt.to(dev), dep.to(dev), N_RGB, net_c, net_f, perturb=0., raw_noise_std=0., **KW)
    worst = max(rel_l2(g, p.grad) for g, p in zip(got["graph"], nets))
    print("  captured graph with in-graph all-reduces vs the union batch on one GPU: worst rel-L2 %.3e" % worst)
    assert worst <= 2e-3

    q = d.FusedQuery(d.get_embedder(10, 0)[0], d.get_embedder(4, 0)[0], 1 << 16, 10, 4, 0)
    rk = dict(network_query_fn=q, perturb=0., N_importance=64, network_fine=net_f, N_samples=64, network_fn=net_c,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0., ndc=True, near=0., far=1.)
    with torch.no_grad():
        ref = d.render_feature_loss(H, W, FOCAL, chunk=1 << 15, rays=(rays[0, :301].to(dev), rays[1, :301].to(dev)),
                                    keep_keys=["rgb_map", "depth_map", "rgb0"], **rk)[-1]
    for k, v in got["patch"].items():
        assert v.shape == ref[k].shape
        err = float((v - ref[k].cpu()).abs().max())
        print("  all-gathered patch %-10s max|diff| vs one GPU %.3e" % (k, err))
        assert err <= 1e-6           # same kernels on the same rays; tile boundaries move with the shard split
