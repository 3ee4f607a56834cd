"""DeviceRayLoader: the iteration protocol of the reference's DataLoader(RayDataset(rays), batch_size, shuffle=True)
(data.py:4-21, run_nerf.py:1126-1206, :1328-1363) with the ray array resident on one device.  Runs on the CPU
device here; the class is device-agnostic torch indexing."""
import numpy as np
import torch

import dlnerf_b200 as dn


def _rays(n):
    return np.arange(n * 9, dtype=np.float32).reshape(n, 3, 3)       # [N, ro+rd+rgb, 3] like rays_rgb


def test_epoch_is_a_permutation_without_replacement():
    data = _rays(1000)
    loader = dn.DeviceRayLoader(data, batch_size=96, device="cpu", generator=torch.Generator().manual_seed(0))
    assert len(loader) == 11
    batches = list(iter(loader))
    assert [b.shape[0] for b in batches] == [96] * 10 + [40] and batches[0].shape[1:] == (3, 3)
    seen = torch.cat(batches)[:, 0, 0].numpy() / 9
    assert sorted(seen.astype(int).tolist()) == list(range(1000))
    assert not np.array_equal(seen, np.arange(1000)), "shuffle=True must permute"
    # a new epoch is a new permutation; the reference restarts the iterator on StopIteration
    it = iter(loader)
    for _ in range(11):
        next(it)
    try:
        next(it)
        raise AssertionError("expected StopIteration at the end of the epoch")
    except StopIteration:
        pass
    again = torch.cat(list(iter(loader)))[:, 0, 0].numpy() / 9
    assert sorted(again.astype(int).tolist()) == list(range(1000)) and not np.array_equal(again, seen)


def test_drop_last_no_shuffle_and_dataset_indexing():
    data = _rays(50)
    loader = dn.DeviceRayLoader(data, batch_size=16, shuffle=False, drop_last=True, device="cpu")
    batches = list(loader)
    assert len(loader) == 3 and len(batches) == 3
    assert torch.equal(torch.cat(batches), torch.from_numpy(data[:48]))
    ds = dn.RayDataset(data, device="cpu")
    assert len(ds) == 50 and torch.equal(ds[7], torch.from_numpy(data[7]))
    sem = np.arange(50)
    ds2 = dn.RayDataset(data, sem, True, device="cpu")
    r, s = ds2[torch.tensor([3, 4])]
    assert r.shape == (2, 3, 3) and s.tolist() == [3, 4]


def test_generator_makes_the_order_reproducible():
    data = _rays(300)
    a = torch.cat(list(dn.DeviceRayLoader(data, 64, device="cpu", generator=torch.Generator().manual_seed(5))))
    b = torch.cat(list(dn.DeviceRayLoader(data, 64, device="cpu", generator=torch.Generator().manual_seed(5))))
    assert torch.equal(a, b)
