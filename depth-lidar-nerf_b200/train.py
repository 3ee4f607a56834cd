"""One training iteration of the hot path, and its ray-sharded data-parallel form.

* ``train_step``  — render (coarse + fine) + RGB / LiDAR-depth loss + backward without building an autograd
  graph: the loss gradient is formed inside the compositing backward kernel
  (``dln_composite_bwd_fused_loss``, north_star part 5) and handed straight to the MLP dgrad / wgrad
  kernels.  Same arithmetic as the drop-in route ``render(...)`` + ``img2mse`` + ``loss.backward()``
  (run_nerf.py:1416-1418, :1451-1466, :1500-1536, :1759-1761, :1773); parameter ``.grad`` tensors are filled
  for the reference's ``optimizer.step()``.
* ``shard_ray_batch`` / ``allreduce_gradients`` — rays are independent, so G GPUs each render an equal
  slice of the RGB rays and of the depth rays (per-class, so the per-rank means are unbiased) and the MLP
  gradients are averaged with ONE all-reduce over a flat buffer.  The reference itself has
  no distributed code; this is the only collective on the path.
"""
from __future__ import annotations

import math
import os as _os
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L
from . import ops
from .run_nerf import FusedQuery
from .run_nerf_helpers import NeRF, ndc_rays

Tensor = torch.Tensor

_DEPTH_MODES = {"mse": 0, "weighted": 1, "weighted_norm": 2, "relative": 3}


# ----------------------------------------------------------------------------------------------------
# ray-sharded data parallelism
# ----------------------------------------------------------------------------------------------------
def shard_bounds(n: int, rank: int, world: int):
    """Contiguous equal split of n items; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_ray_batch(batch_rays: Tensor, target_s: Tensor, target_depth: Optional[Tensor], n_rgb: int, rank: int,
                    world: int, ray_weights: Optional[Tensor] = None, target_semantic: Optional[Tensor] = None):
    """Slice the step's batch [2, n_rgb + n_depth, 3] (RGB rays first, run_nerf.py:1409-1411) for one rank,
    per ray class.  Returns (rays, target_s, target_depth, ray_weights, n_rgb_local); with ``target_semantic`` (one
    class index per RGB ray, run_nerf.py:1331-1332) its slice is appended as a sixth entry."""
    n = batch_rays.shape[1]
    a0, a1 = shard_bounds(n_rgb, rank, world)
    b0, b1 = shard_bounds(n - n_rgb, rank, world)
    rays = torch.cat([batch_rays[:, a0:a1], batch_rays[:, n_rgb + b0:n_rgb + b1]], dim=1)
    td = None if target_depth is None else target_depth[b0:b1]
    rw = None if ray_weights is None else ray_weights[b0:b1]
    if target_semantic is not None:
        return rays, target_s[a0:a1], td, rw, a1 - a0, target_semantic[a0:a1]
    return rays, target_s[a0:a1], td, rw, a1 - a0


def allreduce_gradients(params: Sequence[Tensor], world: int, group=None, average: bool = True) -> None:
    """Average the gradients over the ranks: ONE collective per flat gradient buffer (the kernels hand every
    network's gradients back as views of one fp32 buffer, 2.4 MB for an 8x256 net), issued in place -- on
    NVLink 5 / NVSwitch this is latency-bound, so few large buckets beat many small ones and no staging copy is
    needed.  Only the span of the buffer the parameters' gradients cover is reduced (the kernels' scratch tail behind
    them is rank-local).  Gradients that do not share a buffer (foreign parameters) go through one concatenated
    bucket.  ``average=False``: plain sum (the caller folded 1/world into its loss coefficients)."""
    if world <= 1:
        return
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    bases, loose = {}, []
    for g in grads:
        b = g._base
        if b is not None and b.is_contiguous() and b.dtype == g.dtype and b.dim() == 1:
            lo = g.storage_offset() - b.storage_offset()
            ent = bases.setdefault(id(b), [b, lo, lo + g.numel()])
            ent[1], ent[2] = min(ent[1], lo), max(ent[2], lo + g.numel())
        else:
            loose.append(g)
    for b, lo, hi in bases.values():
        span = b[lo:hi]
        dist.all_reduce(span, op=dist.ReduceOp.SUM, group=group)
        if average:
            span.div_(world)
    if loose:
        flat = torch.cat([g.reshape(-1) for g in loose])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        o = 0
        for g in loose:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()


def render_patch_nograd_sharded(H, W, focal, rays, rank: int, world: int, group=None, keep_keys=None, chunk=1024 * 32,
                                render_fn=None, **kwargs):
    """The no-grad part of a patch render (run_nerf.py:1600-1605: the ~31 k pixels of a 94x352 patch that carry no
    gradient) split over the ranks of a data-parallel job: every rank renders an equal contiguous slice of the rays
    under ``torch.no_grad()`` (forward-only kernels, no stash) and one all_gather per kept map assembles the full
    patch on every rank, in ray order.  The gradient-carrying block of the patch (32x64 rays) is rendered by every
    rank itself, so the patch losses and their gradients are identical on all ranks and the usual gradient average
    leaves them unchanged.  ``rays`` = (rays_o[N,3], rays_d[N,3]); returns {key: [N, ...]} for ``keep_keys``."""
    import torch.distributed as dist
    if render_fn is None:
        from .run_nerf import render_feature_loss as render_fn
    rays_o, rays_d = rays
    n = rays_o.shape[0]
    if n < world:
        raise ValueError("render_patch_nograd_sharded: %d rays cannot be split over %d ranks" % (n, world))
    lo, hi = shard_bounds(n, rank, world)
    with torch.no_grad():
        local = render_fn(H, W, focal, chunk=chunk, rays=(rays_o[lo:hi], rays_d[lo:hi]), keep_keys=keep_keys,
                          **kwargs)[-1]
    if world <= 1:
        return dict(local)
    width = -(-n // world)                       # the largest shard; shorter ones are padded for the collective
    out = {}
    for k in sorted(local):
        v = local[k]
        pad = torch.zeros((width,) + tuple(v.shape[1:]), device=v.device, dtype=v.dtype)
        pad[: hi - lo] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat([parts[r][: shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0]]
                            for r in range(world)], 0)
    return out


# ----------------------------------------------------------------------------------------------------
# fused training step
# ----------------------------------------------------------------------------------------------------
def pack_ray_batch(H: int, W: int, focal: float, batch_rays: Tensor, ndc: bool = True, near: float = 0.,
                   far: float = 1., use_viewdirs: bool = True) -> Tensor:
    """The [N, 8|11] ray_batch of render() (run_nerf.py:145-183): unit view directions from the PRE-warp
    directions, NDC warp with near plane 1, [o, d, near, far, viewdirs]."""
    rays_o, rays_d = batch_rays[0], batch_rays[1]
    if rays_d.is_cuda:
        return ops.pack_rays(H, W, focal, rays_o, rays_d, ndc, near, far, use_viewdirs)
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
    ones = torch.ones_like(rays_d[..., :1])
    cols = [rays_o, rays_d, near * ones, far * ones]
    if viewdirs is not None:
        cols.append(viewdirs)
    return torch.cat(cols, -1).float().contiguous()


def _stash_bytes_per_ray(net: NeRF, n_samples: int) -> int:
    """HBM a ray needs between forward and backward of `net` (activation stashes of both chains + masks)."""
    st = net._state()
    pl = net._plan
    return n_samples * (pl.fwd_slots + pl.bwd_slots) * (L.SLAB_BYTES // L.TILE_ROWS) + n_samples * pl.mask_slots * 32


def default_ray_chunk(network_fn: NeRF, network_fine: NeRF, N_samples: int, N_importance: int,
                      budget_bytes: int = 48 << 30) -> int:
    """Largest multiple of 1024 rays whose stashes fit `budget_bytes` (both nets are live at the same time)."""
    per_ray = _stash_bytes_per_ray(network_fn, N_samples) + _stash_bytes_per_ray(network_fine, N_samples + N_importance)
    return max(1024, (budget_bytes // per_ray) // 1024 * 1024)


_SIDE_STREAMS: Dict[torch.device, "torch.cuda.Stream"] = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    dev = torch.device(dev)
    if dev not in _SIDE_STREAMS:
        _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[dev]


@ops.on_device_of(3)
@torch.no_grad()
def train_step(H, W, focal, batch_rays: Tensor, target_s: Tensor, target_depth: Optional[Tensor], n_rgb: int,
               network_fn: NeRF, network_fine: NeRF, N_samples: int = 64, N_importance: int = 64,
               perturb: float = 1., raw_noise_std: float = 1., white_bkgd: bool = False, lindisp: bool = False,
               ndc: bool = True, near: float = 0., far: float = 1., depth_lambda: float = 0.,
               depth_importance: float = 1., ray_weights: Optional[Tensor] = None, depth_mode: str = "mse",
               coarse_loss: bool = True, world_size: int = 1, group=None, ray_chunk: Optional[int] = None,
               overlap_coarse_backward: bool = True, coarse_sms: Optional[int] = None, fine_sms: Optional[int] = None,
               target_semantic: Optional[Tensor] = None, semantic_lambda: float = 0.,
               rng_state: Optional["ops.RngState"] = None, global_counts=None,
               _rng: Optional[Dict[str, Tensor]] = None, _force_pack: bool = False,
               _coefs: Optional[Tensor] = None, _sem_den: Optional[float] = None) -> Dict[str, Tensor]:
    """render + loss + backward for one ray batch; fills ``.grad`` of both networks (averaged over
    ``world_size`` ranks when > 1) and returns the loss terms as 0-d tensors (no host sync).

    Batches larger than ``ray_chunk`` rays (default: what fits 48 GB of activation stash, ~28 k rays for the
    D=4/D=8 pair) are processed chunk by chunk -- each chunk a slice of the RGB rays plus the matching slice of
    the depth rays, full forward + backward, gradients accumulated in the same flat fp32 buffers -- which is
    the same sum as the unchunked step (config E: N_rand 16 k ... 256 k).

    ``target_semantic`` (class index per RGB ray) with networks that carry the semantic head adds
    ``semantic_lambda * (CE(sem_preds[:n_rgb]) + CE(sem_preds0[:n_rgb]))`` (run_nerf.py:1541-1548): the per-ray logits
    come from one pass over the kept last-trunk-layer activations, the cross-entropy gradient is formed in one small
    kernel and enters the dgrad chain as one fp32 row per ray.

    The reference's four random tensors (rand jitter, randn coarse noise, rand u, randn fine noise) are drawn inside
    the consuming kernels from ``rng_state`` (default: a per-device Philox state seeded from torch's seed); the
    state's device-side counter is advanced in-stream at the end of the step, so a captured graph draws fresh
    numbers on every replay.  ``_rng`` (tests) injects tensors instead.

    ``world_size`` > 1: 1/world_size is folded into the loss-gradient coefficients and the gradients are SUMMED over
    the ranks (one all-reduce per network; the coarse network's is issued from the side stream as soon as its
    backward finishes, under the fine backward).  ``global_counts`` = (n_rgb, n_depth) of the whole batch makes the
    per-rank means exact when the classes do not divide evenly over the ranks.  ``depth_mode='weighted_norm'``
    normalises by max(target_depth) taken on the device (all-reduced MAX over the ranks): no host sync.

    ``_coefs`` (GraphedTrainStep): device fp32[8] = [coef_rgb, coef_depth, depth_norm, depth_lambda*depth_importance,
    coef_rgb(coarse), 0, 1, -] (``loss_coefs``) read by the kernels instead of by-value scalars, so a replayed graph
    follows the depth_importance schedule of run_nerf.py:1527-1532."""
    if network_fine is None or N_importance <= 0:
        raise NotImplementedError("train_step implements the coarse + fine configuration every shipped config uses")
    rb = pack_ray_batch(H, W, focal, batch_rays, ndc, near, far, network_fn.use_viewdirs)
    N, dev = rb.shape[0], rb.device
    network_fn._state(), network_fine._state()          # builds the plans / flat parameter buffers
    n_dep = N - n_rgb
    mode = _DEPTH_MODES[depth_mode]
    use_depth = target_depth is not None and n_dep > 0 and depth_lambda != 0.
    # per-rank mean coefficients; with world_size > 1 the all-reduce SUMS, so 1/world is folded in here -- and with
    # the global class counts the sum over the ranks is exactly the mean over the whole batch
    den_rgb, den_dep = _denominators(n_rgb, n_dep, world_size, global_counts)
    use_sem = target_semantic is not None and semantic_lambda != 0. and n_rgb > 0
    if use_sem and not (network_fn.sem_K and network_fine.sem_K):
        raise RuntimeError("target_semantic needs networks built with semantic_num_classes")
    coef_sem = semantic_lambda / (_sem_den or den_rgb)
    tsem = target_semantic.to(device=dev, dtype=torch.int64) if use_sem else None
    sums = torch.zeros(8, device=dev)
    tgt = ops._f32(target_s, "train_step")
    tdep = ops._f32(target_depth, "train_step") if use_depth else None
    if _coefs is not None:
        coefs = _coefs
    else:
        coefs = torch.tensor(loss_coefs(den_rgb, den_dep, depth_lambda if use_depth else 0., depth_importance,
                                        coarse_loss), dtype=torch.float32).to(dev)
        if use_depth and mode == 2:          # max over the WHOLE batch (run_nerf.py:1518), kept on the device
            coefs[2:3] = batch_depth_max(tdep, world_size, group)
    rw = ops._f32(ray_weights, "train_step") if (use_depth and ray_weights is not None) else None
    S1 = N_samples + N_importance
    if ray_chunk is None:
        ray_chunk = default_ray_chunk(network_fn, network_fine, N_samples, N_importance)
    n_chunks = max(1, -(-N // max(int(ray_chunk), 1)))
    gacc = [torch.zeros(net._plan.n_flat, device=dev, dtype=torch.float32) for net in (network_fn, network_fine)]
    gen = rng_state if rng_state is not None else ops.default_rng(dev, 1)
    main = torch.cuda.current_stream(dev)
    advanced = []

    for c in range(n_chunks):
        if n_chunks == 1:
            rb_c, tgt_c, tdep_c, rw_c, nr_c, rng = rb, tgt, tdep, rw, n_rgb, (_rng or {})
            tsem_c = tsem
        else:
            r0, r1 = shard_bounds(n_rgb, c, n_chunks)
            d0, d1 = shard_bounds(n_dep, c, n_chunks)
            pick = lambda t: torch.cat([t[r0:r1], t[n_rgb + d0:n_rgb + d1]], 0)     # noqa: E731
            rb_c, tgt_c, nr_c = pick(rb), tgt[r0:r1], r1 - r0
            tdep_c = tdep[d0:d1] if tdep is not None else None
            rw_c = rw[d0:d1] if rw is not None else None
            tsem_c = tsem[r0:r1] if tsem is not None else None
            rng = {k: pick(v) for k, v in (_rng or {}).items()}
            if rb_c.shape[0] == 0:
                continue
        Nc = rb_c.shape[0]
        rays_d = rb_c[:, 3:6].contiguous()

        def draw(name, k):
            """(injected tensor | None, in-kernel generator reference | None): tensor k of chunk c"""
            if name in rng:
                return rng[name], None
            return None, (gen, 4 * c + k)

        t_rand, g_t = draw("t_rand", 0) if perturb > 0. else (None, None)
        sem_kw = lambda S: dict(sem_group=S) if use_sem else {}          # noqa: E731
        pack_ev = None
        if c == 0 and overlap_coarse_backward and not _os.environ.get("DLN_NO_PACK_OVERLAP"):
            # the fine network's fp32 -> bf16 weight re-pack (fold + two pack launches, ~40 us) does not depend on
            # anything of this step: it runs on the side stream under the coarse forward
            s0 = _side_stream(dev)
            s0.wait_stream(main)
            with torch.cuda.stream(s0):
                network_fine._pack(network_fine._state(), force=_force_pack)
                pack_ev = torch.cuda.Event()
                pack_ev.record(s0)
        if t_rand is None and network_fn.fused_sampling_available():
            # stratified sampling fused into the coarse chain's tile prologue (north_star part 1): the kernel computes
            # the depths it encodes and writes z0 for the compositing kernels -- no separate launch, no z round trip
            z0 = torch.empty(Nc, N_samples, device=dev, dtype=torch.float32)
            mode0 = ("rays", dict(rng=g_t, lindisp=lindisp))
        else:
            z0 = ops.stratified_z(rb_c, N_samples, t_rand, lindisp, rng=g_t)
            mode0 = "rays"
        raw0, saved0, *sem0 = network_fn._run_forward(mode0, rb_c, z0, Nc * N_samples, keep=True,
                                                      force_pack=_force_pack and c == 0, **sem_kw(N_samples))
        raw0 = raw0.view(Nc, N_samples, -1)
        noise0, g_n0 = draw("noise0", 1) if raw_noise_std > 0. else (None, None)
        u, g_u = draw("u", 2) if perturb != 0. else (None, None)
        if N_samples == 64 and 1 <= N_importance <= 64 and not _os.environ.get("DLN_NO_FUSED_RESAMPLE"):
            # raw2outputs of the coarse pass and the resampling of its weights in one launch (both are launch-latency
            # class at 4096 rays; same bits as the two calls below)
            rgb0, disp0, acc0, w0, depth0, z_samples, z1 = ops.composite_resample(
                raw0, z0, rays_d, noise0, float(raw_noise_std), bool(white_bkgd),
                N_importance, u, rng_noise=g_n0, rng_u=g_u)
        else:
            rgb0, disp0, acc0, w0, depth0 = ops.composite(raw0, z0, rays_d, noise0, float(raw_noise_std),
                                                          bool(white_bkgd), rng=g_n0)
            z_samples, z1 = ops.importance_resample(z0, w0, N_importance, u, rng=g_u)
        side = None
        grads_c = None

        def coarse_backward():
            # coarse pass: colour loss only (run_nerf.py:1759-1761); depth_map0 is unsupervised
            d_raw0 = ops.composite_bwd_fused_loss(raw0, z0, rays_d, noise0, raw_noise_std, white_bkgd, tgt_c,
                                                  None, None, nr_c, 0., 0., 0, 1., sums[2:4], rng=g_n0,
                                                  coefs_dev=coefs[4:7])
            # the coarse logits are supervised too (run_nerf.py:1545-1546), whatever no_coarse says
            d_sem0 = ops.semantic_ce(sem0[0], tsem_c, nr_c, coef_sem, sums[5:6]) if use_sem else None
            cap = coarse_sms if side is not main else None
            g = network_fn._run_backward(d_raw0, saved0, Nc * N_samples, gflat=gacc[0], sms=cap, wgrad_sms=cap, d_sem=d_sem0)
            if world_size > 1 and c == n_chunks - 1:
                # the coarse network's gradients are complete: reduce them from the side stream, under the fine
                # network's kernels still running on the main stream
                _assign_grads(network_fn, g)
                allreduce_gradients(list(network_fn.parameters()), world_size, group, average=False)
            if c == n_chunks - 1 and side is not main and not _rng and not early:
                # every kernel that draws from the generator's current base has been enqueued (the main stream's before
                # this stream forked): the counter bump rides on the side stream, off the step's critical path
                gen.advance(4 * n_chunks)
                advanced.append(True)
            return g

        early = bool(coarse_loss or use_sem) and overlap_coarse_backward and coarse_sms is not None
        if early:
            # SM-partitioned schedule: the coarse backward depends on the coarse forward alone, and the fine network's
            # forward and dgrad kernels are bound by their stash WRITES (3.9 TB/s for a pure write stream), not by the
            # SMs.  So they run on `fine_sms` SMs, and the coarse backward -- its wgrad a pure READ stream, which the
            # memory system overlaps with writes almost for free -- runs next to them on the remaining `coarse_sms`.
            side = _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                grads_c = coarse_backward()
        if pack_ev is not None:
            main.wait_event(pack_ev)
        raw1, saved1, *sem1 = network_fine._run_forward("rays", rb_c, z1, Nc * S1, keep=True,
                                                        force_pack=_force_pack and c == 0 and pack_ev is None,
                                                        sms=fine_sms if early else None, **sem_kw(S1))
        raw1 = raw1.view(Nc, S1, -1)
        noise1, g_n1 = draw("noise1", 3) if raw_noise_std > 0. else (None, None)
        # fine pass: colour loss on the RGB rays, depth loss on the depth rays (run_nerf.py:1461, :1500-1524)
        d_raw1 = ops.composite_bwd_fused_loss(raw1, z1, rays_d, noise1, raw_noise_std, white_bkgd, tgt_c, tdep_c,
                                              rw_c, nr_c, 0., 0., mode, 1., sums[0:2], rng=g_n1, coefs_dev=coefs[0:3])
        d_sem1 = ops.semantic_ce(sem1[0], tsem_c, nr_c, coef_sem, sums[4:5]) if use_sem else None
        if (coarse_loss or use_sem) and not early:
            # default schedule: the coarse backward on a second stream next to the fine backward -- its CTAs fill the
            # SMs the persistent fine-net kernels leave idle at their tails and under the HBM-bound wgrad
            side = _side_stream(dev) if overlap_coarse_backward else main
            if side is not main:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                grads_c = coarse_backward()
        grads_f = network_fine._run_backward(d_raw1, saved1, Nc * S1, gflat=gacc[1], d_sem=d_sem1,
                                             sms=fine_sms if early else None)
        if world_size > 1 and c == n_chunks - 1:
            _assign_grads(network_fine, grads_f)
            allreduce_gradients(list(network_fine.parameters()), world_size, group, average=False)
        if side is not None and side is not main:
            main.wait_stream(side)
        del saved1, d_raw1, raw1, saved0
    if not _rng and not advanced:
        gen.advance(4 * n_chunks)             # the next step (or graph replay) names fresh tensors
    _assign_grads(network_fine, grads_f)
    if coarse_loss or use_sem:
        _assign_grads(network_fn, grads_c)
    img_loss = sums[0] / (3.0 * max(n_rgb, 1))
    img_loss0 = sums[2] / (3.0 * max(n_rgb, 1))
    depth_loss = sums[1] / max(n_dep, 1)
    loss = img_loss + coefs[3] * depth_loss + (img_loss0 if coarse_loss else 0.)
    out = {"img_loss": img_loss, "img_loss0": img_loss0, "depth_loss": depth_loss, "psnr": -10. * torch.log10(img_loss)}
    if use_sem:
        out["semantic_loss"], out["semantic_loss0"] = sums[4] / n_rgb, sums[5] / n_rgb
        loss = loss + semantic_lambda * (out["semantic_loss"] + out["semantic_loss0"])
    out["loss"] = loss
    return out


def _denominators(n_rgb: int, n_dep: int, world_size: int, global_counts):
    """Divisors of the two loss means as the gradient coefficients see them: the all-reduce SUMS over the ranks, so
    they are the whole batch's class counts (given, or local count x world size)."""
    if global_counts and world_size > 1:
        return float(max(global_counts[0], 1)), float(max(global_counts[1], 1))
    return float(max(n_rgb, 1) * world_size), float(max(n_dep, 1) * world_size)


def loss_coefs(den_rgb: float, den_dep: float, depth_lambda: float, depth_importance: float, coarse_loss: bool):
    """The 8 floats the fused-loss kernels read from device memory: fine pass [0:3] = d(mean sq. colour error) /
    d(colour) scale 2/(3 n_rgb), depth scale 2 lambda importance / n_depth (run_nerf.py:1500-1536), depth_norm
    (patched on the device for 'weighted_norm'); [3] = the depth term's weight in the reported loss; coarse pass
    [4:7] = colour only (:1759-1761)."""
    c_rgb = 2.0 / (3.0 * den_rgb)
    return [c_rgb, 2.0 * depth_lambda * depth_importance / den_dep, 1.0, depth_lambda * depth_importance,
            c_rgb if coarse_loss else 0.0, 0.0, 1.0, 0.0]


def batch_depth_max(target_depth: Tensor, world_size: int = 1, group=None) -> Tensor:
    """max(target_depth) over the whole batch (run_nerf.py:1518) as a 1-element device tensor; all ranks' shards
    contribute (all-reduce MAX).  No host sync."""
    m = torch.amax(target_depth.reshape(-1), 0, keepdim=True).float()
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return m


def _assign_grads(net: NeRF, grads) -> None:
    for p, g in zip(net._ordered_params(), grads):
        if g is not None:
            p.grad = g


class GraphedTrainStep:
    """``train_step`` captured once into a CUDA graph and replayed per iteration (static shapes: N_rand is fixed
    in the reference's loop).  The ~20 launches of a step (our kernels, the weight re-pack, memsets, the generator's
    counter bump) then cost one graph launch; the fp32 -> bf16 weight re-pack is part of the graph, so an
    ``optimizer.step()`` between replays is picked up; the random tensors are drawn in-kernel from a device-side
    counter the graph itself advances, so every replay sees fresh numbers.  The loss coefficients live in a small
    device tensor: ``__call__(..., depth_importance=...)`` follows the reference's per-iteration decay
    (run_nerf.py:1527-1532), and ``depth_mode='weighted_norm'`` takes max(target_depth) on the device before the
    replay.  With ``world_size`` > 1 the two gradient all-reduces are captured in the graph (NCCL supports stream
    capture), the coarse network's under the fine backward; ``capture_allreduce=False`` runs them after the replay.

        step = GraphedTrainStep(H, W, focal, n_rays, n_rgb, net_c, net_f, depth_lambda=0.01, ...)
        out = step(batch_rays, target_s, target_depth)        # fills .grad, returns 0-d loss tensors
    """

    def __init__(self, H, W, focal, n_rays: int, n_rgb: int, network_fn: NeRF, network_fine: NeRF,
                 world_size: int = 1, group=None, warmup: int = 3, capture_allreduce: bool = True,
                 global_counts=None, **kw):
        dev = next(network_fn.parameters()).device
        self.world_size, self.group = world_size, group
        self.capture_allreduce = bool(capture_allreduce) and world_size > 1
        self.nets = (network_fn, network_fine)
        self.rays = torch.zeros(2, n_rays, 3, device=dev)
        self.rays[1, :, 2] = -1.0                                  # any valid direction for the warm-up passes
        self.target_s = torch.zeros(n_rgb, 3, device=dev)
        self.target_depth = torch.ones(n_rays - n_rgb, device=dev)
        self.ray_weights = torch.ones(n_rays - n_rgb, device=dev) if kw.pop("use_ray_weights", False) else None
        self.target_semantic = (torch.zeros(n_rgb, device=dev, dtype=torch.int64)
                                if kw.get("semantic_lambda", 0.) != 0. else None)
        self.rng = kw.pop("rng_state", None) or ops.RngState(dev, ops.default_rng(dev, 1).seed)
        self._norm = kw.get("depth_mode") == "weighted_norm"
        self._lambda = float(kw.get("depth_lambda", 0.))
        self._importance = float(kw.pop("depth_importance", 1.))
        self._coarse = bool(kw.get("coarse_loss", True))
        ws = world_size if (self.capture_allreduce or world_size == 1) else 1
        self._den = _denominators(n_rgb, n_rays - n_rgb, world_size, global_counts)
        self.coefs = torch.tensor(loss_coefs(*self._den, self._lambda, self._importance, self._coarse),
                                  dtype=torch.float32).to(dev)
        args = (H, W, focal, self.rays, self.target_s, self.target_depth, n_rgb, network_fn, network_fine)
        kw = dict(kw, ray_weights=self.ray_weights, target_semantic=self.target_semantic, world_size=ws, group=group,
                  global_counts=global_counts if ws > 1 else None, _force_pack=True, _coefs=self.coefs,
                  rng_state=self.rng)
        if ws == 1 and world_size > 1:
            # all-reduce after the replay: the kernels still see the whole-batch divisors through `coefs`
            kw["_sem_den"] = self._den[0]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                train_step(*args, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = train_step(*args, **kw)
        self._grads = [(p, p.grad) for net in self.nets for p in net.parameters()]

    def close(self) -> None:
        """Drop the captured graph.  With ``world_size`` > 1 the graph holds NCCL kernels: destroy it BEFORE
        ``dist.destroy_process_group()`` (a process group torn down under a live graph that references its communicator
        hung at exit on the 2-GPU box)."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def set_depth_importance(self, depth_importance: float) -> None:
        """run_nerf.py:1527-1532: the decayed weight of the depth term, for the following replays (two scalar
        fills on the stream, no host sync)."""
        if depth_importance == self._importance:
            return
        self._importance = float(depth_importance)
        c = loss_coefs(*self._den, self._lambda, self._importance, self._coarse)
        self.coefs[1:2].fill_(c[1])
        self.coefs[3:4].fill_(c[3])

    def __call__(self, batch_rays: Tensor, target_s: Tensor, target_depth: Optional[Tensor] = None,
                 ray_weights: Optional[Tensor] = None, target_semantic: Optional[Tensor] = None,
                 depth_importance: Optional[float] = None) -> Dict[str, Tensor]:
        self.rays.copy_(batch_rays, non_blocking=True)
        self.target_s.copy_(target_s, non_blocking=True)
        if target_depth is not None:
            self.target_depth.copy_(target_depth, non_blocking=True)
        if ray_weights is not None and self.ray_weights is not None:
            self.ray_weights.copy_(ray_weights, non_blocking=True)
        if target_semantic is not None and self.target_semantic is not None:
            self.target_semantic.copy_(target_semantic, non_blocking=True)
        if depth_importance is not None:
            self.set_depth_importance(depth_importance)
        if self._norm:
            self.coefs[2:3] = batch_depth_max(self.target_depth, self.world_size, self.group)
        self.graph.replay()
        for p, g in self._grads:          # survive optimizer.zero_grad(set_to_none=True)
            p.grad = g
        if self.world_size > 1 and not self.capture_allreduce:
            allreduce_gradients([p for p, _ in self._grads], self.world_size, self.group, average=False)
        return self.out
