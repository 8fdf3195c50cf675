"""CPU tests of the "next" rows: PLY I/O with the reference's field names and volume -> cloud construction."""
import os
import sys

import numpy as np
import pytest
import torch

from point_unet_b200 import data_prepare as dp
from point_unet_b200.helper_ply import read_ply, write_ply

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference/PointSegment"


def _cloud(n=257, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32), rng.standard_normal((n, 4)).astype(np.float32),
            rng.integers(0, 4, n).astype(np.uint8))


def test_ply_round_trip_brats_fields(tmp_path):
    xyz, colors, labels = _cloud()
    path = str(tmp_path / "case")
    assert write_ply(path, (xyz, colors, labels), ["x", "y", "z", "t1ce", "t1", "flair", "t2", "class"])
    data = read_ply(path + ".ply")
    assert data.dtype.names == ("x", "y", "z", "t1ce", "t1", "flair", "t2", "class")
    assert np.array_equal(np.vstack((data["x"], data["y"], data["z"])).T, xyz)           # runBraTS.py:100
    assert np.array_equal(np.vstack((data["t1ce"], data["t1"], data["flair"], data["t2"])).T, colors)
    assert np.array_equal(data["class"], labels) and data["class"].dtype == np.uint8
    assert not write_ply(path, (xyz, labels[:-1]), ["x", "y", "z", "class"])            # inconsistent lengths -> False


def test_reads_file_written_by_the_reference_writer():
    """tests/golden/ref_written.ply was produced by the reference's own write_ply (generator in this test, run where
    /root/reference exists); our reader must parse it, and our writer must emit the same bytes."""
    fixture = os.path.join(GOLD, "ref_written.ply")
    rng = np.random.default_rng(42)
    xyz = rng.random((33, 3), dtype=np.float32)
    value = rng.standard_normal((33, 1)).astype(np.float32)
    cls = rng.integers(0, 2, 33).astype(np.uint8)
    if os.path.isdir(REF) and not os.path.exists(fixture):
        sys.path.insert(0, REF)
        import helper_ply as ref_ply  # the reference module (plain numpy)
        ref_ply.write_ply(fixture, (xyz, value, cls), ["x", "y", "z", "value", "class"])
        sys.path.remove(REF)
    if not os.path.exists(fixture):
        pytest.skip("fixture not generated")
    data = read_ply(fixture)
    assert np.array_equal(np.vstack((data["x"], data["y"], data["z"])).T, xyz)
    assert np.array_equal(data["value"], value[:, 0]) and np.array_equal(data["class"], cls)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        write_ply(os.path.join(d, "ours"), (xyz, value, cls), ["x", "y", "z", "value", "class"])
        assert open(os.path.join(d, "ours.ply"), "rb").read() == open(fixture, "rb").read()


def test_pancreas_cloud_construction():
    g = torch.Generator().manual_seed(0)
    shape = (24, 20, 12)
    img = torch.randn(shape, generator=g) * 50 + 100
    label = torch.zeros(shape, dtype=torch.uint8)
    label[8:14, 6:12, 3:8] = 1
    n_fg = int(label.sum())
    z = dp.zscore_volume(img, nonzero_only=False)
    assert abs(float(z.mean())) < 1e-9 and abs(float(z.std(unbiased=False)) - 1) < 1e-9
    c = dp.sample_pancreas_cloud(z, label, n_point=1000, generator=g)
    assert c["xyz"].shape == (1000, 3) and c["xyz"].dtype == torch.float32
    assert bool((c["labels"][:n_fg] == 1).all()) and bool((c["labels"][n_fg:] == 0).all())   # foreground first, unshuffled
    vox = c["xyz_origin"].long()
    assert len({tuple(v) for v in vox.tolist()}) == 1000                                      # without replacement
    want = vox.to(torch.float32) / torch.tensor(shape, dtype=torch.float32)
    assert torch.equal(c["xyz"], want)
    fg = torch.nonzero(label > 0)
    assert torch.equal(vox[:n_fg], fg)                                                        # x-major order of the loops


def test_brats_cloud_construction():
    g = torch.Generator().manual_seed(1)
    shape = (20, 18, 14)
    mods = torch.rand((4,) + shape, generator=g) * 100
    brain = torch.zeros(shape, dtype=torch.bool)
    brain[3:17, 2:16, 2:12] = True
    mods = mods * brain
    label = torch.zeros(shape, dtype=torch.uint8)
    label[8:12, 7:11, 5:9] = 2
    z = torch.stack([dp.zscore_volume(m, nonzero_only=True) for m in mods])
    assert bool((z[:, ~brain] == 0).all())
    c = dp.sample_brats_cloud(z.float(), label, num_points=1500, generator=g)
    assert c["xyz"].shape == (1500, 3) and c["colors"].shape == (1500, 4)
    assert int((c["labels"] > 0).sum()) == int((label > 0).sum())                             # every tumour voxel is kept
    assert len(set(c["point_idx"].tolist())) == 1500
    vox = c["xyz_origin_all"][c["point_idx"].long()]
    want = (vox.double() / torch.tensor(shape, dtype=torch.float64)).float()
    assert torch.equal(c["xyz"], want)
    assert not torch.equal(c["labels"], torch.sort(c["labels"], descending=True).values)      # shuffled, not tumour-first
