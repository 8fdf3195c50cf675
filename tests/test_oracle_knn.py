"""CPU tests pinning the KNN oracle (oracle/knn_oracle.c) to the reference.

Pins: (1) the committed golden vectors, which are verbatim outputs of the reference's own compiled C++
(tests/golden/make_knn_golden.py); (2) when oracle/_ref is present (authoring container), live runs of the
reference on fresh seeds.  Also checks the canonical (distance, index) rule against exhaustive search and
its relation to nanoflann's visiting-order rule (SURVEY.md section 8c "tie rule").
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import knn as ok
from point_unet_b200 import synthetic as syn
from tests.golden.make_knn_golden import SMALL, make_cloud

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLD, "knn_golden.npz"))


@pytest.fixture(scope="module")
def checksums():
    with open(os.path.join(GOLD, "knn_checksums.json")) as f:
        return json.load(f)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("kind,n,seed", SMALL)
def test_restatement_matches_reference_golden(golden, kind, n, seed):
    tag = f"{kind}_{n}"
    p = golden[tag + "_xyz"]
    # inputs regenerate bit-identically from the seed
    assert np.array_equal(p, make_cloud(kind, n, seed)[None])
    assert np.array_equal(ok.knn_restated(p, p, 16, tie_rule=0), golden[tag + "_k16_self"])
    assert np.array_equal(ok.knn_restated(p[:, : n // 4], p, 1, tie_rule=0), golden[tag + "_k1_prefix"])


def test_restatement_ragged_and_tiny(golden):
    t = golden["tiny_xyz"]
    assert np.array_equal(ok.knn_restated(t, t, 16, tie_rule=0), golden["tiny_k16_self"])
    assert (golden["tiny_k16_self"][:, :, 10:] == 0).all()  # N1 < K: trailing slots stay zero (knn.pyx:93)
    r = golden["ragged_xyz"]
    assert np.array_equal(ok.knn_restated(r, r, 16, tie_rule=0), golden["ragged_k16_self"])
    assert np.array_equal(ok.knn_restated(r[:, :351], r, 1, tie_rule=0), golden["ragged_k1_prefix"])
    assert np.array_equal(ok.knn_restated(r[:, :500], r[:, 100:], 5, tie_rule=0), golden["ragged_k5_cross"])


@pytest.mark.parametrize("tag", ["uniform_16384", "uniform_65536", "jitter_65536", "pancreas_65536", "pancreas_180000"])
def test_restatement_matches_reference_checksums(checksums, tag):
    c = checksums[tag]
    p = make_cloud(c["kind"], c["n"], c["seed"])[None]
    assert digest(p) == c["xyz"]
    r16 = ok.knn_restated(p, p, 16, tie_rule=0)
    assert digest(r16) == c["k16_self_idx"]
    sub = p[:, : c["n"] // 4]
    r1 = ok.knn_restated(sub, p, 1, tie_rule=0)
    assert digest(r1) == c["k1_prefix_idx"]
    assert digest(ok.knn_dists(p, p, r16)) == c["k16_self_dist"]


@pytest.mark.skipif(not ok.have_reference(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("kind", ["uniform", "jitter", "pancreas", "brats"])
def test_restatement_matches_live_reference(kind):
    for seed in (11, 12):
        p = np.stack([make_cloud(kind, 20000, seed * 10 + b) for b in range(2)])
        for k in (1, 3, 16):
            assert np.array_equal(ok.knn_reference(p, p, k), ok.knn_restated(p, p, k, tie_rule=0))
        assert np.array_equal(ok.knn_reference(p[:, :5000], p, 1), ok.knn_restated(p[:, :5000], p, 1, tie_rule=0))


@pytest.mark.parametrize("kind", ["uniform", "pancreas", "brats"])
def test_canonical_rule_is_exhaustive_search(kind):
    p = make_cloud(kind, 3000, 21)[None]
    bi, bd = ok.knn_brute(p, p, 16, return_dist=True)
    ci, cd = ok.knn_restated(p, p, 16, tie_rule=1, return_dist=True)
    assert np.array_equal(bi, ci) and np.array_equal(bd, cd)
    b1 = ok.knn_brute(p[:, :750], p, 1)
    assert np.array_equal(b1, ok.knn_restated(p[:, :750], p, 1, tie_rule=1))


def test_canonical_vs_nanoflann_rule(golden):
    """Tie-free clouds: identical rows.  Lattice clouds: identical distance rows, indices differ only in ties."""
    for kind, n, _ in SMALL:
        tag = f"{kind}_{n}"
        p = golden[tag + "_xyz"]
        ref = golden[tag + "_k16_self"]
        can, cd = ok.knn_restated(p, p, 16, tie_rule=1, return_dist=True)
        rd = ok.knn_dists(p, p, ref)
        assert np.array_equal(rd, cd), tag  # sorted K-distance vectors identical in 100 % of rows
        if kind in ("uniform", "jitter"):
            assert np.array_equal(ref, can), tag
        else:
            # rows without a tie inside the top K+1 are identical
            c17, d17 = ok.knn_restated(p, p, 17, tie_rule=1, return_dist=True)
            tie_free = (np.diff(d17, axis=-1) != 0).all(-1)
            assert np.array_equal(ref[tie_free], can[tie_free]), tag
            assert tie_free.mean() < 1.0  # the lattice really does tie


def test_synthetic_cloud_shapes():
    c = syn.pancreas_cloud(20000, 3)
    assert c["xyz"].shape == (20000, 3) and c["xyz"].dtype == np.float32 and c["features"].shape == (20000, 1)
    assert c["labels"][: c["labels"].sum()].all()  # foreground first, unshuffled
    assert len(np.unique(c["xyz_origin"].astype(np.int64) @ np.array([1 << 40, 1 << 20, 1]))) == 20000
    b = syn.brats_cloud(20000, 3)
    assert b["features"].shape == (20000, 4) and set(np.unique(b["labels"])) <= {0, 1, 2, 3}
    assert (b["xyz"] >= 0).all() and (b["xyz"] < 1).all()
