"""Volume -> point cloud construction ("context-aware sampling").

Device-agnostic torch code (runs on the GPU when the volume lives there) restating what the reference's offline
preparation does, so a dense volume can be turned into the network's input without leaving the device:

* Pancreas (``utils/dataPreparePancreas.py:31-46,132-169``): z-score the whole CT volume; take ALL voxels with
  label > 0 first (x-major order), then ``n_point - #foreground`` background voxels drawn without replacement;
  ``xyz_origin = voxel (uint16)``, ``xyz = voxel.astype(f32) / shape.astype(f32)``.
* BraTS (``utils/dataPrepareBraTS.py:32-116`` + ``runBraTS.py:100-119``): z-score each modality over its non-zero
  voxels (zeros stay 0); brain = voxels where any modality is non-zero; take all tumour voxels plus random
  non-tumour brain voxels up to ``num_points``, then shuffle; ``xyz = (voxel_f64 / shape).astype(f32)``.

The reference draws the background with Python's ``random.sample`` -- results are equal in distribution, not bit for
bit; everything deterministic (ordering, normalisation arithmetic, dtypes) follows the reference exactly.
"""
from __future__ import annotations

import torch


def zscore_volume(volume: torch.Tensor, nonzero_only: bool) -> torch.Tensor:
    """``(v - mean) / std`` with population std (numpy's default).  ``nonzero_only``: statistics over v > 0 and zeros stay
    zero (BraTS, dataPrepareBraTS.py:34-52); otherwise over the whole volume (Pancreas, dataPreparePancreas.py:32-46)."""
    v = volume.double()
    if nonzero_only:
        m = v > 0
        pix = v[m]
        out = (v - pix.mean()) / pix.std(unbiased=False)
        return torch.where(volume == 0, torch.zeros_like(out), out)
    return (v - v.mean()) / v.std(unbiased=False)


def _voxel_coords(mask: torch.Tensor) -> torch.Tensor:
    """Integer coordinates [M,3] of the set voxels in x-major (C) order, like the reference's nested x,y,z loops."""
    return torch.nonzero(mask)


def sample_pancreas_cloud(img: torch.Tensor, label: torch.Tensor, n_point: int = 180000, generator=None,
                          background_choice=None):
    """dict(xyz f32 [N,3], value f32 [N,1], labels uint8 [N], xyz_origin int16-range [N,3]); foreground first, unshuffled.
    ``background_choice`` (optional int64 [n_point - #foreground]): positions among the background voxels in x-major order
    to take instead of a fresh random draw (what the reference's ``random.sample(none_tumor, k)`` returns)."""
    shape = torch.tensor(img.shape, device=img.device)
    fg = _voxel_coords(label > 0)
    n_bg = n_point - fg.shape[0]
    if n_bg < 0:
        raise ValueError(f"{fg.shape[0]} foreground voxels exceed n_point={n_point}")
    bg_all = _voxel_coords(label == 0)
    if background_choice is not None:
        pick = torch.as_tensor(background_choice, dtype=torch.long, device=img.device)
        assert pick.numel() == n_bg
    else:
        pick = torch.randperm(bg_all.shape[0], device=img.device, generator=generator)[:n_bg]
    vox = torch.cat([fg, bg_all[pick]], dim=0)
    xyz = vox.to(torch.float32) / shape.to(torch.float32)          # dataPreparePancreas.py:163 (fp32 division)
    value = img[vox[:, 0], vox[:, 1], vox[:, 2]].to(torch.float32).unsqueeze(1)
    labels = label[vox[:, 0], vox[:, 1], vox[:, 2]].to(torch.uint8)
    return dict(xyz=xyz, value=value, labels=labels, xyz_origin=vox.to(torch.int32))


def sample_brats_cloud(modalities: torch.Tensor, label: torch.Tensor, num_points: int = 180000, generator=None,
                       background_choice=None, shuffle_perm=None):
    """``modalities [4,X,Y,Z]`` (already z-scored), ``label [X,Y,Z]`` -> dict(xyz f32 [N,3], colors f32 [N,4], labels,
    point_idx (row of each sampled point among ALL brain points -- the ``p_idx`` of testBraTS.py:226-231), xyz_origin_all).
    ``background_choice`` / ``shuffle_perm`` (optional): the positions among the non-tumour brain points and the final
    permutation to use instead of fresh random draws (``random.sample`` / ``DP.shuffle_idx`` in runBraTS.py:110-114)."""
    shape = torch.tensor(label.shape, device=label.device)
    brain = (modalities != 0).any(dim=0)                            # dataPrepareBraTS.py:78
    vox_all = _voxel_coords(brain)
    lab_all = label[vox_all[:, 0], vox_all[:, 1], vox_all[:, 2]]
    tumor = torch.nonzero(lab_all > 0).squeeze(1)
    none_tumor = torch.nonzero(lab_all == 0).squeeze(1)
    n_bg = num_points - tumor.shape[0]
    if n_bg < 0 or n_bg > none_tumor.shape[0]:
        raise ValueError("cannot draw the requested number of points from this volume")
    if background_choice is not None:
        pick = none_tumor[torch.as_tensor(background_choice, dtype=torch.long, device=label.device)]
        assert pick.numel() == n_bg
    else:
        pick = none_tumor[torch.randperm(none_tumor.shape[0], device=label.device, generator=generator)[:n_bg]]
    idx = torch.cat([tumor, pick])
    perm = torch.as_tensor(shuffle_perm, dtype=torch.long, device=label.device) if shuffle_perm is not None \
        else torch.randperm(idx.shape[0], device=label.device, generator=generator)
    idx = idx[perm]                                                                       # DP.shuffle_idx, runBraTS.py:114
    vox = vox_all[idx]
    xyz = (vox.double() / shape.double()).to(torch.float32)         # dataPrepareBraTS.py:85-89 (fp64 division, then cast)
    colors = modalities[:, vox[:, 0], vox[:, 1], vox[:, 2]].t().to(torch.float32).contiguous()
    return dict(xyz=xyz, colors=colors, labels=lab_all[idx].to(torch.uint8), point_idx=idx.to(torch.int32),
                xyz_origin_all=vox_all.to(torch.int32))


def save_xyz_origin(path: str, xyz_origin: torch.Tensor, dataset: str) -> None:
    """``<ID>_xyz_origin_loop_<i>.npy`` (Pancreas: uint16, dataPreparePancreas.py:160-161) / ``<ID>_xyz_origin.npy`` (BraTS:
    platform int, dataPrepareBraTS.py:81-82) -- the integer voxel coordinates test mode scatters through (testPancreas.py:193)."""
    import numpy as np
    a = xyz_origin.detach().cpu().numpy()
    np.save(path, a.astype(np.uint16) if dataset == "Pancreas" else a.astype(int))


def load_xyz_origin(path: str, device="cpu") -> torch.Tensor:
    """Read either flavour back as int32 ``[n,3]`` (x, y, z) for ``ops.point2prod`` / ``ops.point2label``."""
    import numpy as np
    return torch.from_numpy(np.load(path).astype(np.int32)).to(device)
