"""Drop-in for the hot-path names of the reference's ``run_nerf.py``:

    batchify, run_network, batchify_rays, render, render_rays, create_nerf

Same signatures and return structures (run_nerf.py:50-194, :389-675).  ``render_rays`` takes the fused
B200 route (stratified depths -> fused sample/encode/MLP kernel -> compositing -> CDF inversion + merge ->
fine pass) whenever the networks are this package's ``NeRF`` and ``network_query_fn`` is the one
``create_nerf`` built; any other callable goes through the reference's generic composition with each
stage still running on this library's kernels.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .run_nerf_helpers import NeRF, get_embedder, ndc_rays, raw2outputs, sample_pdf, to8b

Tensor = torch.Tensor


def batchify(fn, chunk):
    """run_nerf.py:50-57."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_nerf.py:60-74 (generic route: encode, concatenate, apply ``fn`` in netchunk slices)."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded = torch.cat([embedded, embeddirs_fn(input_dirs_flat)], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class FusedQuery:
    """``network_query_fn`` built by create_nerf.  Calling it behaves exactly like the reference's lambda
    (run_nerf.py:434-437); render_rays recognises the type and skips materialising pts / encodings."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk, multires, multires_views, i_embed):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk
        self.L_pts = 0 if i_embed == -1 else multires
        self.L_dir = 0 if i_embed == -1 else multires_views

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn, embeddirs_fn=self.embeddirs_fn,
                           netchunk=self.netchunk)

    def fused_ok(self, net, ray_batch) -> bool:
        if not isinstance(net, NeRF):
            return False
        want_dir = 3 + 6 * self.L_dir if net.use_viewdirs else net.input_ch_views
        return (net.input_ch == 3 + 6 * self.L_pts and net.input_ch_views == want_dir
                and (not net.use_viewdirs or ray_batch.shape[-1] > 9))


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_nerf.py:77-89.  The fused kernels have no activation-memory reason to chunk, but the chunk loop
    is kept because the reference draws its random numbers per chunk."""
    all_ret: Dict[str, list] = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def _ray_batch(H, W, focal, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam=None, depths=None):
    """The flat [N, 8 | 9 | 11 | 12] batch render_rays consumes (run_nerf.py:138-183) and the leading shape of the
    rays: [o, d, near, far, (depth), (unit view direction taken BEFORE the NDC warp)].  Device tensors with scalar
    near / far -- the training loop -- are packed by one kernel (ops.pack_rays); everything else (static camera,
    per-ray depths, tensor-valued bounds, CPU tensors) goes through the equivalent torch operations."""
    origins, dirs = get_rays(H, W, focal, c2w) if c2w is not None else rays
    lead = list(dirs.shape[:-1])
    scalar_bounds = not torch.is_tensor(near) and not torch.is_tensor(far)
    if c2w_staticcam is None and depths is None and scalar_bounds and torch.is_tensor(dirs) and dirs.is_cuda:
        return ops.pack_rays(H, W, focal, origins, dirs, ndc, near, far, use_viewdirs), lead
    unit = None
    if use_viewdirs:
        unit = (dirs / dirs.norm(dim=-1, keepdim=True)).reshape(-1, 3).float()
        if c2w_staticcam is not None:          # view directions from c2w, geometry from the static camera (:150-152)
            origins, dirs = get_rays(H, W, focal, c2w_staticcam)
    if ndc:
        origins, dirs = ndc_rays(H, W, focal, 1., origins, dirs)
    origins, dirs = origins.reshape(-1, 3).float(), dirs.reshape(-1, 3).float()
    ones = torch.ones_like(dirs[:, :1])
    parts = [origins, dirs, near * ones, far * ones]
    if depths is not None:
        parts.append(depths.reshape(-1, 1))
    if unit is not None:
        parts.append(unit)
    return torch.cat(parts, -1), lead


def _unflatten(all_ret, lead):
    return {k: v.reshape(lead + list(v.shape[1:])) for k, v in all_ret.items()}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, depths=None, **kwargs):
    """run_nerf.py:112-194.  Returns [rgb_map, disp_map, acc_map, depth_map, extras]."""
    packed, lead = _ray_batch(H, W, focal, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam, depths)
    out = _unflatten(batchify_rays(packed, chunk, **kwargs), lead)
    head = ('rgb_map', 'disp_map', 'acc_map', 'depth_map')
    return [out[k] for k in head] + [{k: v for k, v in out.items() if k not in head}]


def batchify_rays_feature_loss(rays_flat, chunk=1024 * 32, keep_keys=None, **kwargs):
    """run_nerf.py:90-108: batchify_rays that only keeps the entries named in ``keep_keys`` (patch renders)."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            if keep_keys and k not in keep_keys:
                continue
            all_ret.setdefault(k, []).append(ret[k])
    return {k: torch.cat(all_ret[k], 0) for k in all_ret}


def render_feature_loss(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1.,
                        use_viewdirs=False, c2w_staticcam=None, keep_keys=None, **kwargs):
    """run_nerf.py:197-265: the patch render of the feature / GAN / inverse-depth losses (:1552-1647).  The caller
    renders the few gradient-carrying rays of a patch normally and the rest under ``torch.no_grad()``; without an
    autograd graph the MLP kernels run in their forward-only form (no activation stash, no ReLU masks).  Returns
    ``[rgb_map, disp_map, acc_map (those kept)] + [dict of every kept entry]``."""
    packed, lead = _ray_batch(H, W, focal, rays, c2w, ndc, near, far, use_viewdirs, c2w_staticcam)
    out = _unflatten(batchify_rays_feature_loss(packed, chunk, keep_keys=keep_keys, **kwargs), lead)
    head = [k for k in ('rgb_map', 'disp_map', 'acc_map') if not keep_keys or k in keep_keys]
    return [out[k] for k in head] + [dict(out)]


def render_path(render_poses, hwf, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0, iteration=0,
                writer=None, coords=None):
    """run_nerf.py:268-359: full-image renders for a list of poses (evaluation / video), no autograd graph and
    therefore no activation stash.  Returns (rgbs[P,H,W,3], disps[P,H,W]) as numpy like the reference; with
    ``savedir`` each view's maps go to ``{i:03d}.npz`` (+ an 8-bit PNG when cv2 is importable -- imageio, the
    reference's writer, is not a dependency here).  With ``semantic_loss`` in ``render_kwargs`` a third array is
    returned like the reference does (:355-357) -- here the per-pixel class indices argmax(sem_preds) [P,H,W]; the
    colour palette (SemanticSegmentorHelper) and the Tensorboard ``writer`` images are the caller's visualisation
    code and are not reproduced."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    H, W = int(H), int(W)
    rgbs, disps, sems = [], [], []
    for i, c2w in enumerate(render_poses):
        with torch.no_grad():
            rgb, disp, acc, depth, extras = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], retraw=True,
                                                   **render_kwargs)
        rgbs.append(rgb.cpu().numpy())
        disps.append(disp.cpu().numpy())
        if 'sem_preds' in extras:
            sems.append(torch.argmax(extras['sem_preds'], dim=-1).cpu().numpy())
        if savedir is not None:
            rgb8 = to8b(np.nan_to_num(rgbs[-1]))
            np.savez(os.path.join(savedir, '{:03d}.npz'.format(i)), rgb=rgbs[-1], disp=disps[-1],
                     acc=acc.cpu().numpy(), depth=depth.cpu().numpy())
            try:
                import cv2
                cv2.imwrite(os.path.join(savedir, '{:03d}.png'.format(i)), rgb8[..., ::-1])
            except ImportError:
                pass
    if sems:
        return np.stack(rgbs, 0), np.stack(disps, 0), np.stack(sems, 0)
    return np.stack(rgbs, 0), np.stack(disps, 0)


def get_rays(H, W, focal, c2w):
    """run_nerf_helpers.py:266-282 (full-image rays for render_path; host-side set-up, torch ops)."""
    dev = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev),
                          indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


@ops.on_device_of(0)
def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                sigma_loss=None, semantic_loss=False, _rng=None):
    """run_nerf.py:520-675.  The reference's four random tensors -- rand[N,S] (jitter, :585), randn[N,S] (coarse
    density noise, helpers:565), rand[N,Ni] (u, helpers:509), randn[N,S+Ni] -- are drawn INSIDE the consuming kernels
    (Philox4x32-10 keyed by torch's seed; tensor k of call c is named by the offset 4c + k, and the compositing
    backward regenerates its noise from the same name).  ``_rng`` (private, tests only) injects tensors instead:
    dict(t_rand, noise0, u, noise1); ``_rng['state']`` = an ``ops.RngState`` to draw from."""
    if sigma_loss is not None:
        raise NotImplementedError("sigma_loss reads an undefined variable in the reference train loop "
                                  "(run_nerf.py:1527) and is not part of the hot path")
    if network_fn is None:
        raise NotImplementedError("the alpha_model branch (run_nerf.py:606-622) is dead code in the reference")
    rb = ops._f32(ray_batch, "render_rays")
    N = rb.shape[0]
    dev = rb.device
    rays_o, rays_d = rb[:, 0:3], rb[:, 3:6]
    viewdirs = rb[:, -3:] if rb.shape[-1] > 9 else None
    rng = _rng or {}
    gen = rng.get("state") or ops.default_rng(dev, 0)
    off = gen.next_offsets(4)

    def draw(name, k):
        """(injected tensor | None, in-kernel generator reference | None)"""
        if name in rng:
            return rng[name], None
        return None, (gen, off + k)

    def query(net, z, strat=None):
        """(raw for compositing, per-ray semantic logits | None, raw as the reference returns it).  Fused route: the
        logits come from the kept activations of the last trunk layer (one sum per ray, csrc/semantic_kernels.cu);
        compositing runs on the 4-channel raw and the per-sample logits are appended only to the returned copy.
        Generic route: raw[N, S, 4+K] from the query function, summed over the samples (helpers:589)."""
        if isinstance(network_query_fn, FusedQuery) and network_query_fn.fused_ok(net, rb):
            K = net.sem_K
            if semantic_loss and not K:
                raise RuntimeError("semantic_loss=True needs networks built with semantic_num_classes")
            if not (semantic_loss or (retraw and K)):
                raw = net.forward_rays(rb, z, strat=strat)
                return raw, None, raw
            raw, sem, pts = net.forward_rays(rb, z, semantic=bool(semantic_loss), point_logits=bool(retraw), strat=strat)
            return raw, sem, (torch.cat([raw, pts], -1) if retraw else raw)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
        raw = network_query_fn(pts, viewdirs, net)
        return raw, (ops.sample_sum(raw, 4) if semantic_loss else None), raw

    def composite(raw, z, noise_name, k):
        noise, g = draw(noise_name, k) if raw_noise_std > 0. else (None, None)
        return ops.composite(raw, z, rays_d, noise, float(raw_noise_std), bool(white_bkgd), rng=g)

    t_rand, g = draw("t_rand", 0) if perturb > 0. else (None, None)
    if (t_rand is None and isinstance(network_query_fn, FusedQuery) and network_query_fn.fused_ok(network_fn, rb)
            and network_fn.fused_sampling_available()):
        # stratified sampling fused into the coarse network's kernel: it computes the depths it encodes and fills z_vals
        z_vals = torch.empty(N, N_samples, device=dev, dtype=torch.float32)
        raw, sem, raw_ret = query(network_fn, z_vals, strat=dict(rng=g, lindisp=lindisp))
    else:
        z_vals = ops.stratified_z(rb, N_samples, t_rand, lindisp, rng=g)
        raw, sem, raw_ret = query(network_fn, z_vals)
    rgb_map, disp_map, acc_map, weights, depth_map = composite(raw, z_vals, "noise0", 1)

    ret = {}
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0, depth_map0, sem0 = rgb_map, disp_map, acc_map, depth_map, sem
        u, g = draw("u", 2) if perturb != 0. else (None, None)                  # det = (perturb == 0)
        z_samples, z_vals = ops.importance_resample(z_vals, weights.detach(), N_importance, u, rng=g)
        run_fn = network_fn if network_fine is None else network_fine
        raw, sem, raw_ret = query(run_fn, z_vals)
        rgb_map, disp_map, acc_map, weights, depth_map = composite(raw, z_vals, "noise1", 3)
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map, depth_map=depth_map)
    if retraw:
        ret['raw'] = raw_ret
    if semantic_loss:
        ret['sem_preds'] = sem                                                     # :652-653
    if N_importance > 0:
        ret['rgb0'], ret['disp0'], ret['acc0'], ret['depth_map0'] = rgb_map_0, disp_map_0, acc_map_0, depth_map0
        ret['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)
        if semantic_loss:
            ret['sem_preds0'] = sem0                                               # :662-663
    return ret


def create_nerf(args):
    """run_nerf.py:389-517: embedders, coarse + fine NeRF, query fn, Adam, checkpoint reload, kwargs dicts.
    Differences: no torchsummary printout (:511-515, crashes when model_fine is None), and the alpha_model
    branch (:405-419) is rejected."""
    device = torch.device("cuda")
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    if getattr(args, "alpha_model_path", None) is not None:
        raise NotImplementedError("alpha_model_path: the two-stage variant is not part of the hot path")
    sem = getattr(args, "semantic_num_classes", None)
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                 input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs, semantic_num_classes=sem).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs,
                          semantic_num_classes=sem).to(device)
        grad_vars += list(model_fine.parameters())
    network_query_fn = FusedQuery(embed_fn, embeddirs_fn, args.netchunk, args.multires,
                                  getattr(args, "multires_views", 0), args.i_embed)
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    basedir, expname = args.basedir, args.expname
    if getattr(args, "ft_path", None) is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        d = os.path.join(basedir, expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'tar' in f] if os.path.isdir(d) else []
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt['global_step']
        if not args.no_reload_optimizer:
            optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        for net, key in ((model, 'network_fn_state_dict'), (model_fine, 'network_fine_state_dict')):
            if net is None:
                continue
            cur = net.state_dict()
            cur.update({k: v for k, v in ckpt[key].items() if k in cur})      # key-filtered merge (:466-477)
            net.load_state_dict(cur)

    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model,
        'use_viewdirs': args.use_viewdirs, 'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std,
        'semantic_loss': getattr(args, "semantic_loss", False),
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    else:
        render_kwargs_train['ndc'] = True
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
