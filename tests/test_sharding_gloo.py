"""CPU, world_size 2, gloo: the ray-sharded data-parallel plumbing (shard_ray_batch + allreduce_gradients).
The per-rank loss uses the reference's normalisation (means over the local RGB rays / local depth rays);
averaging the per-rank gradients must reproduce the single-process gradient of the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dlnerf_b200 as dn


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))


def _loss(model, rays, tgt, dep, n_rgb, lam=0.3):
    x = torch.cat([rays[0], rays[1]], -1)            # [N, 6]
    y = model(x)
    return torch.mean((y[:n_rgb, :3] - tgt) ** 2) + lam * torch.mean((y[n_rgb:, 3] - dep) ** 2)


def _batch(n_rgb=12, n_dep=8):
    g = torch.Generator().manual_seed(1)
    return torch.randn(2, n_rgb + n_dep, 3, generator=g), torch.rand(n_rgb, 3, generator=g), torch.rand(n_dep, generator=g)


def _worker(rank, world, port, out, flat_views=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rays, tgt, dep = _batch()
    n_rgb = tgt.shape[0]
    r, t, d, _, n_loc = dn.shard_ray_batch(rays, tgt, dep, n_rgb, rank, world)
    model = _model()
    _loss(model, r, t, d, n_loc).backward()
    if flat_views:      # the layout the kernels produce: every .grad a view of one flat buffer (+ a scratch tail)
        ps = list(model.parameters())
        flat = torch.cat([p.grad.reshape(-1) for p in ps] + [torch.full((5,), float(rank))])
        o = 0
        for p in ps:
            p.grad = flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        extra = torch.nn.Parameter(torch.zeros(3))          # a foreign parameter with its own gradient tensor
        extra.grad = torch.full((3,), float(rank + 1))
        dn.allreduce_gradients(ps + [extra], world)
        # the foreign gradient is averaged; the scratch tail behind the parameters' span is rank-local and untouched
        assert torch.allclose(extra.grad, torch.full((3,), 1.5)) and torch.equal(flat[-5:], torch.full((5,), float(rank)))
        probe = torch.nn.Parameter(torch.zeros(2))
        buf = torch.full((4,), float(rank + 1))
        probe.grad = buf[1:3]
        dn.allreduce_gradients([probe], world, average=False)       # plain sum (1/world folded into the loss scale)
        assert torch.equal(buf, torch.tensor([rank + 1.0, 3.0, 3.0, rank + 1.0]))
    else:
        dn.allreduce_gradients(list(model.parameters()), world)
    if rank == 0:
        torch.save([p.grad for p in model.parameters()], out)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            spans = [dn.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@pytest.mark.parametrize("flat_views", [False, True])
def test_sharded_gradients_match_single_process(tmp_path, flat_views):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, port, out, flat_views), nprocs=2, join=True)
    got = torch.load(out)
    rays, tgt, dep = _batch()
    model = _model()
    _loss(model, rays, tgt, dep, tgt.shape[0]).backward()
    for g, p in zip(got, model.parameters()):
        torch.testing.assert_close(g, p.grad, rtol=1e-5, atol=1e-6)


def test_shard_keeps_class_order():
    rays, tgt, dep = _batch(n_rgb=10, n_dep=6)
    seen_rgb, seen_dep = [], []
    for rank in range(4):
        r, t, d, _, n_loc = dn.shard_ray_batch(rays, tgt, dep, 10, rank, 4)
        assert r.shape[1] == n_loc + d.shape[0] and t.shape[0] == n_loc
        seen_rgb.append(r[:, :n_loc])
        seen_dep.append(r[:, n_loc:])
    assert torch.equal(torch.cat(seen_rgb, 1), rays[:, :10]) and torch.equal(torch.cat(seen_dep, 1), rays[:, 10:])
    # semantic targets (one class index per RGB ray) follow the RGB slice
    tsem = torch.arange(10)
    parts = [dn.shard_ray_batch(rays, tgt, dep, 10, rank, 4, target_semantic=tsem) for rank in range(4)]
    assert all(len(p) == 6 and p[5].shape[0] == p[4] for p in parts)
    assert torch.equal(torch.cat([p[5] for p in parts]), tsem)


def _fake_patch_render(H, W, focal, chunk, rays, keep_keys=None, **kw):
    """Stands in for render_feature_loss (GPU only): per-ray maps that identify the ray."""
    o, d = rays
    out = {"rgb_map": o * 2.0 + d, "depth_map": o[:, 0] - d[:, 2], "acc_map": d[:, 1]}
    return [{k: v for k, v in out.items() if not keep_keys or k in keep_keys}]


def _patch_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(4)
    rays = (torch.randn(37, 3, generator=g), torch.randn(37, 3, generator=g))        # ragged: 19 + 18 rays
    got = dn.render_patch_nograd_sharded(94, 352, 138.14, rays, rank, world, keep_keys=["rgb_map", "depth_map"],
                                         render_fn=_fake_patch_render)
    if rank == 1:
        torch.save(got, out)
    dist.destroy_process_group()


def test_sharded_patch_render_gathers_the_full_patch_in_ray_order(tmp_path):
    """SURVEY 8(f) rank 3, multi-GPU part: the no-grad rays of a patch are split over the ranks and all-gathered."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "patch.pt")
    mp.spawn(_patch_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(4)
    rays = (torch.randn(37, 3, generator=g), torch.randn(37, 3, generator=g))
    ref = _fake_patch_render(94, 352, 138.14, 1, rays, keep_keys=["rgb_map", "depth_map"])[-1]
    assert set(got) == {"rgb_map", "depth_map"}
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
