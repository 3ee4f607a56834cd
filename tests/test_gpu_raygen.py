"""GPU parity tests (through the C ABI) of the device ray generators: get_rays_np,
get_rays_by_coord_np, get_rays_cropped_feature_loss_new and the two training-bank builders, against the vectors
the UNMODIFIED reference produced (tests/golden/raygen.npz, oracle/make_golden.py) -- bit for bit -- and against the
oracle at the LLFF / KITTI-360 image sizes."""
import os

import numpy as np
import pytest
import torch

from gpu_util import O, dn

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "raygen.npz"))


def _hwf(g):
    return int(g["HWf"][0]), int(g["HWf"][1]), float(g["HWf"][2])


def test_get_rays_np_matches_the_reference_bit_for_bit(g):
    H, W, focal = _hwf(g)
    for n in range(3):
        o, d = dn().get_rays_np(H, W, focal, g["poses"][n])
        assert o.is_cuda and o.shape == (H, W, 3) and d.dtype == torch.float32
        np.testing.assert_array_equal(d.cpu().numpy(), g["grid_d%d" % n])
        np.testing.assert_array_equal(o.cpu().numpy(), np.broadcast_to(g["poses"][n][:, 3], (H, W, 3)))
    # a stack of poses in one launch == the reference's list comprehension (run_nerf.py:1126)
    o, d = dn().get_rays_np(H, W, focal, torch.from_numpy(g["poses"]).to(DEV))
    assert d.shape == (3, H, W, 3)
    for n in range(3):
        np.testing.assert_array_equal(d[n].cpu().numpy(), g["grid_d%d" % n])


@pytest.mark.parametrize("tag", ["64", "32"])
def test_get_rays_by_coord_np_matches_the_reference_bit_for_bit(g, tag):
    H, W, focal = _hwf(g)
    o, d = dn().get_rays_by_coord_np(H, W, focal, g["poses"][1], g["coord" + tag])
    assert d.dtype == (torch.float64 if tag == "64" else torch.float32)
    np.testing.assert_array_equal(d.cpu().numpy(), g["coord_d" + tag])
    np.testing.assert_array_equal(o.cpu().numpy(), g["coord_o" + tag])
    # empty input (an image without LiDAR points)
    o, d = dn().get_rays_by_coord_np(H, W, focal, g["poses"][1], np.zeros((0, 2)))
    assert o.shape == (0, 3) and d.shape == (0, 3)


def test_cropped_patch_rays_match_the_reference_bit_for_bit(g):
    H, W, focal = _hwf(g)
    for n in range(int(g["n_crops"][0])):
        nH, nW, gH, gW, sw, sh = (int(v) for v in g["crop%d_cfg" % n])
        grad, nograd, crop = dn().get_rays_cropped_feature_loss_new(
            H, W, focal, torch.from_numpy(g["poses"][2]).to(DEV), nH=nH, nW=nW, gradH=gH, gradW=gW, _start=(sw, sh),
            _perm=torch.from_numpy(g["crop%d_perm" % n]))
        assert crop == [sw, sw + nW - 1, sh, sh + nH - 1]
        assert grad[0].shape == (gH * gW, 3) and nograd[1].shape == (nH * nW - gH * gW, 3)
        assert grad[2].dtype == torch.int64
        np.testing.assert_array_equal(torch.cat([grad[1], nograd[1]]).cpu().numpy(), g["crop%d_d" % n])
        np.testing.assert_array_equal(torch.cat([grad[2], nograd[2]]).cpu().numpy(), g["crop%d_pts" % n])
        pose_o = np.broadcast_to(g["poses"][2][:, 3], (nH * nW, 3))
        np.testing.assert_array_equal(torch.cat([grad[0], nograd[0]]).cpu().numpy(), pose_o)


def test_cropped_patch_rays_draw_like_the_reference():
    """Live randomness: the crop corner consumes numpy's generator exactly as the reference does (two randint draws),
    the split is a permutation of the crop, and grad + no-grad pixels tile the crop once."""
    H, W, focal = 94, 352, 138.2        # KITTI-360 at the reference's training resolution
    pose = torch.eye(4)[:3].to(DEV)
    np.random.seed(5)
    grad, nograd, crop = dn().get_rays_cropped_feature_loss_new(H, W, focal, pose, nH=32, nW=32, gradH=4, gradW=4)
    np.random.seed(5)
    sw, sh = np.random.randint(0, W - 32 + 1), np.random.randint(0, H - 32 + 1)
    assert crop == [sw, sw + 31, sh, sh + 31]
    pts = torch.cat([grad[2], nograd[2]]).cpu()
    flat = (pts[:, 0] * 32 + pts[:, 1]).sort().values
    assert torch.equal(flat, torch.arange(32 * 32))
    # every ray passes through its pixel: d = ((x - W/2)/f, -(y - H/2)/f, -1) for the identity pose
    d = torch.cat([grad[1], nograd[1]]).cpu()
    x = (sw + pts[:, 1]).float()
    y = (sh + pts[:, 0]).float()
    assert torch.equal(d[:, 0], (x - W * .5) / np.float32(focal)) and torch.equal(d[:, 1], -((y - H * .5) / np.float32(focal)))


def test_full_size_banks_match_the_oracle():
    """Config-B image size (378 x 504, 20 views) and a KITTI-360-like depth set: the device-built banks equal the
    reference's numpy construction (run_nerf.py:1126-1187, restated inline) before shuffling; the shuffle is a
    permutation of the rows."""
    H, W, focal = 378, 504, 407.56
    rs = np.random.RandomState(0)
    n_img = 5
    poses = np.stack([np.concatenate([np.linalg.qr(rs.randn(3, 3))[0], rs.randn(3, 2)], 1) for _ in range(n_img)]).astype(np.float32)
    images = rs.rand(n_img, H, W, 3).astype(np.float32)
    i_train = np.array([0, 2, 3])
    # reference construction (numpy, :1126-1147)
    rays = np.stack([np.stack(O.get_rays_np(H, W, focal, p[:3, :4]), 0) for p in poses], 0)
    ref = np.concatenate([rays, images[:, None]], 1).transpose(0, 2, 3, 1, 4)
    ref = np.stack([ref[i] for i in i_train], 0).reshape(-1, 3, 3).astype(np.float32)
    bank = dn().build_ray_bank(H, W, focal, poses, images, i_train, shuffle=False)
    assert bank.is_cuda and bank.shape == ref.shape
    np.testing.assert_array_equal(bank.cpu().numpy(), ref)
    gen = torch.Generator(device="cuda").manual_seed(3)
    shuffled = dn().build_ray_bank(H, W, focal, poses, images, i_train, shuffle=True, generator=gen)
    key = lambda b: torch.unique(b.reshape(-1, 9), dim=0)
    assert shuffled.shape == bank.shape and not torch.equal(shuffled, bank) and torch.equal(key(shuffled), key(bank))

    depth_gts = [{"coord": np.stack([rs.uniform(0, W, 700 + 13 * i), rs.uniform(0, H, 700 + 13 * i)], -1),
                  "depth": rs.uniform(2, 60, 700 + 13 * i), "weight": rs.uniform(0, 1, 700 + 13 * i)} for i in range(n_img)]
    rows = []
    for i in i_train:          # :1170-1177
        rd = np.stack(O.get_rays_by_coord_np(H, W, focal, poses[i, :3, :4], depth_gts[i]["coord"]), 0).transpose(1, 0, 2)
        dv = np.repeat(depth_gts[i]["depth"][:, None, None], 3, axis=2)
        wv = np.repeat(depth_gts[i]["weight"][:, None, None], 3, axis=2)
        rows.append(np.concatenate([rd, dv, wv], 1))
    ref_d = np.concatenate(rows, 0).astype(np.float32)
    bank_d, max_depth = dn().build_depth_ray_bank(H, W, focal, poses, depth_gts, i_train, shuffle=False)
    np.testing.assert_array_equal(bank_d.cpu().numpy(), ref_d)
    assert float(max_depth) == float(ref_d[:, 3, 0].max())


def test_bank_with_labels_feeds_the_device_loader():
    """run_nerf.py:1153-1158 + :1195-1206: per-pixel labels follow the shuffle of the bank, and the device loader hands out
    (rays, labels) batches of it."""
    H, W, focal = 24, 40, 31.5
    rs = np.random.RandomState(1)
    poses = np.stack([np.concatenate([np.linalg.qr(rs.randn(3, 3))[0], rs.randn(3, 1)], 1) for _ in range(3)]).astype(np.float32)
    images = rs.rand(3, H, W, 3).astype(np.float32)
    seg = rs.randint(0, 19, (3, H, W))
    bank, labels = dn().build_ray_bank(H, W, focal, poses, images, [0, 2], shuffle=True, segmentation_gt=seg,
                                       generator=torch.Generator(device="cuda").manual_seed(0))
    assert bank.shape == (2 * H * W, 3, 3) and labels.shape == (2 * H * W,)
    # a row's colour identifies its pixel: the label of that pixel must have travelled with it
    flat_img = torch.from_numpy(images[[0, 2]].reshape(-1, 3)).to(DEV)
    flat_seg = torch.from_numpy(seg[[0, 2]].reshape(-1)).to(DEV)
    idx = (bank[:, 2][:, None, :] == flat_img[None, :64, :]).all(-1).float().argmax(0)      # rows holding the first 64 pixels
    assert torch.equal(labels[idx], flat_seg[:64])
    loader = dn().DeviceRayLoader(bank, batch_size=500, semantic_data=labels, device=DEV)
    rays, lab = next(iter(loader))
    assert rays.shape == (500, 3, 3) and lab.shape == (500,) and rays.is_cuda
