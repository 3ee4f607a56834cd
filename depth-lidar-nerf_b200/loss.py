"""Mirror of the reference's loss.py for the one loss that sits directly on the rendered patch:
``InverseDepthSmoothnessLoss`` (loss.py:55-133; instantiated as ``depth_inv_loss``, run_nerf.py:1249, applied to the
accumulated patch depth / colour at :1646).  Same class name, argument checks and messages; the arithmetic is one
forward and one backward kernel (``dln_inv_depth_smooth_fwd/bwd``) instead of ~65 element-wise launches."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class _InvDepthSmooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idepth, image):
        d, im = ops._f32(idepth, "InverseDepthSmoothnessLoss"), ops._f32(image, "InverseDepthSmoothnessLoss")
        N, _, H, W = d.shape
        sums = torch.zeros(2, device=d.device)
        L.call("dln_inv_depth_smooth_fwd", d.data_ptr(), im.data_ptr(), N, H, W, sums.data_ptr(), ops._stream(),
               tag="inv_depth_smooth_fwd")
        ctx.save_for_backward(d, im)
        cx, cy = N * H * (W - 1), N * (H - 1) * W
        # torch.mean of an empty tensor is NaN in the reference (W == 1 or H == 1)
        lx = sums[0] / cx if cx > 0 else sums[0] * float("nan")
        ly = sums[1] / cy if cy > 0 else sums[1] * float("nan")
        return lx + ly

    @staticmethod
    def backward(ctx, g):
        d, im = ctx.saved_tensors
        N, _, H, W = d.shape
        gd = torch.empty_like(d) if ctx.needs_input_grad[0] else None
        gi = torch.empty_like(im) if ctx.needs_input_grad[1] else None
        if gd is None and gi is None:
            return None, None
        gs = g.reshape(1).float().contiguous()
        L.call("dln_inv_depth_smooth_bwd", d.data_ptr(), im.data_ptr(), N, H, W, gs.data_ptr(),
               None if gd is None else gd.data_ptr(), None if gi is None else gi.data_ptr(), ops._stream(),
               tag="inv_depth_smooth_bwd")
        return gd, gi


class InverseDepthSmoothnessLoss(nn.Module):
    r"""loss = |d_x d_ij| exp(-||d_x I_ij||) + |d_y d_ij| exp(-||d_y I_ij||) (means over the patch), loss.py:55-133.
    Inverse depth ``(N, 1, H, W)``, image ``(N, 3, H, W)``, output scalar."""

    def forward(self, idepth: torch.Tensor, image: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(idepth):
            raise TypeError("Input idepth type is not a torch.Tensor. Got {}".format(type(idepth)))
        if not torch.is_tensor(image):
            raise TypeError("Input image type is not a torch.Tensor. Got {}".format(type(image)))
        if not len(idepth.shape) == 4:
            raise ValueError("Invalid idepth shape, we expect BxCxHxW. Got: {}".format(idepth.shape))
        if not len(image.shape) == 4:
            raise ValueError("Invalid image shape, we expect BxCxHxW. Got: {}".format(image.shape))
        if not idepth.shape[-2:] == image.shape[-2:]:
            raise ValueError("idepth and image shapes must be the same. Got: {}".format(idepth.shape, image.shape))
        if not idepth.device == image.device:
            raise ValueError("idepth and image must be in the same device. Got: {}".format(idepth.device, image.device))
        if not idepth.dtype == image.dtype:
            raise ValueError("idepth and image must be in the same dtype. Got: {}".format(idepth.dtype, image.dtype))
        if idepth.shape[1] != 1 or image.shape[1] != 3 or idepth.shape[0] != image.shape[0]:
            raise NotImplementedError("the kernel handles the reference's use: idepth (N,1,H,W), image (N,3,H,W)")
        return _InvDepthSmooth.apply(idepth, image)
