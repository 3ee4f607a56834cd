"""depth-lidar-nerf_b200 — B200 (sm_100a) implementation of the ray-rendering training hot path of
mertkiray/depth-lidar-nerf behind the reference's own Python call signatures.

The directory name carries a hyphen, so import it as ``import dlnerf_b200`` (alias module at the repo
root) or ``importlib.import_module("depth-lidar-nerf_b200")``.

Layout:  csrc/ (CUDA kernels + C ABI, include/dlnerf_b200.h)  ·  _lib.py (loader)  ·  plan.py (static MLP
plans)  ·  ops.py (torch wrappers)  ·  run_nerf_helpers.py / run_nerf.py (mirrors of the reference modules)
·  train.py (fused train step, ray-sharded data parallel)  ·  optim.py (flat Adam)  ·  data.py (device ray bank).
"""
from . import _lib
from ._lib import build, lib
from . import ops
from .run_nerf_helpers import (Embedder, NeRF, get_embedder, get_rays_by_coord_np, get_rays_cropped_feature_loss_new,
                               get_rays_np, img2mse, mse2psnr, ndc_rays, raw2outputs, sample_pdf, to8b)
from .run_nerf import (FusedQuery, batchify, batchify_rays, batchify_rays_feature_loss, create_nerf, get_rays,
                       render, render_feature_loss, render_path, render_rays, run_network)
from .data import DeviceRayLoader, RayDataset, build_depth_ray_bank, build_ray_bank
from .optim import FlatAdam
from .loss import InverseDepthSmoothnessLoss
from .train import (GraphedTrainStep, allreduce_gradients, default_ray_chunk, pack_ray_batch,
                    render_patch_nograd_sharded, shard_bounds, shard_ray_batch, train_step)

__all__ = ["build", "lib", "ops", "Embedder", "NeRF", "get_embedder", "img2mse", "mse2psnr", "ndc_rays",
           "raw2outputs", "sample_pdf", "to8b", "FusedQuery", "batchify", "batchify_rays", "create_nerf",
           "get_rays", "render", "render_rays", "run_network", "allreduce_gradients", "pack_ray_batch",
           "shard_bounds", "shard_ray_batch", "train_step", "GraphedTrainStep", "default_ray_chunk", "FlatAdam",
           "render_path", "DeviceRayLoader", "RayDataset", "render_feature_loss", "batchify_rays_feature_loss", "InverseDepthSmoothnessLoss",
           "render_patch_nograd_sharded", "get_rays_np", "get_rays_by_coord_np", "get_rays_cropped_feature_loss_new",
           "build_ray_bank", "build_depth_ray_bank"]
