"""Device-resident ray bank and batch sampler.

The reference keeps all training rays in a numpy array, wraps it in ``RayDataset`` (data.py:4-21: one
``torch.Tensor(self.rayData[index])`` per RAY) and draws batches with
``iter(DataLoader(RayDataset(rays), batch_size=N, shuffle=True, num_workers=0))``, restarting the iterator on
``StopIteration`` (run_nerf.py:1126-1206, :1328-1363).  At 1 M rays/s that per-ray Python work would be the
bottleneck long before the GPU is, so here the ray array is uploaded once and a batch is one ``index_select`` on
the device; the iteration protocol (epoch = one random permutation without replacement, ``StopIteration`` at its
end, last batch short unless ``drop_last``) is the DataLoader's.

    rays = dn.DeviceRayLoader(rays_rgb, batch_size=N_rgb)          # replaces RayDataset + DataLoader
    it = iter(rays)
    try: batch = next(it)
    except StopIteration: it = iter(rays); batch = next(it)        # the reference's own restart idiom
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

Tensor = torch.Tensor


class RayDataset(torch.utils.data.Dataset):
    """data.py:4-21 with the array held as a tensor on ``device`` (indexing returns device tensors)."""

    def __init__(self, ray_data, semantic_data=None, use_semantic_data=False, device=None):
        super().__init__()
        self.rayData = torch.as_tensor(np.asarray(ray_data) if not torch.is_tensor(ray_data) else ray_data,
                                       dtype=torch.float32)
        if device is not None:
            self.rayData = self.rayData.to(device)
        self.length = self.rayData.shape[0]
        self.use_semantic_data = use_semantic_data
        if use_semantic_data:
            self.semantic_data = torch.as_tensor(semantic_data).to(self.rayData.device)

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        if self.use_semantic_data:
            return self.rayData[index], self.semantic_data[index]
        return self.rayData[index]


class DeviceRayLoader:
    """Batches of rays gathered on the device; iteration protocol of DataLoader(shuffle=True, num_workers=0)."""

    def __init__(self, ray_data, batch_size: int, shuffle: bool = True, drop_last: bool = False, device="cuda",
                 generator: Optional[torch.Generator] = None, semantic_data=None):
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        self.dataset = RayDataset(ray_data, semantic_data, semantic_data is not None, device=device)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator = generator

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        n, dev = len(self.dataset), self.dataset.rayData.device
        if self.shuffle:
            g = self.generator
            gdev = g.device if g is not None else dev
            order = torch.randperm(n, device=gdev, generator=g).to(dev)
        else:
            order = torch.arange(n, device=dev)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            yield self.dataset[idx]
