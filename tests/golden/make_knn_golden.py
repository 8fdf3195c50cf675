"""Generates tests/golden/knn_golden.npz + knn_checksums.json from the REFERENCE's own compiled C++
(oracle/_ref/libknn_ref.so, built from /root/reference by `make -C oracle ref`).  Run in the authoring
container only:  python tests/golden/make_knn_golden.py

Small cases store inputs and the reference's int32 output verbatim.  Large cases store sha256 digests of
(a) the reference index rows (meaningful for tie-free clouds) and (b) the fp32 distance rows of the
reference's neighbours (identical under any tie rule), the inputs being regenerated from seeds.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import knn as ok  # noqa: E402
from point_unet_b200 import synthetic as syn  # noqa: E402


def make_cloud(kind, n, seed):
    if kind == "uniform":
        return syn.uniform_cloud(n, seed)
    if kind == "jitter":
        return syn.jittered_lattice_cloud(n, seed)
    if kind == "pancreas":
        return syn.pancreas_cloud(n, seed)["xyz"]
    if kind == "brats":
        return syn.brats_cloud(n, seed)["xyz"]
    raise ValueError(kind)


SMALL = [(kind, n, seed) for kind in ("uniform", "jitter", "pancreas", "brats") for n, seed in ((1000, 1), (4096, 2))]
LARGE = [("uniform", 16384, 16384), ("uniform", 65536, 65536), ("jitter", 65536, 5), ("pancreas", 65536, 6),
         ("uniform", 180000, 180000), ("jitter", 180000, 7), ("pancreas", 180000, 0), ("brats", 180000, 0)]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert ok.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    arrays, sums = {}, {}
    for kind, n, seed in SMALL:
        p = make_cloud(kind, n, seed)[None]
        tag = f"{kind}_{n}"
        arrays[tag + "_xyz"] = p
        arrays[tag + "_k16_self"] = ok.knn_reference(p, p, 16)
        sub = p[:, : n // 4]
        arrays[tag + "_k1_prefix"] = ok.knn_reference(sub, p, 1)
    # a ragged / tiny family: N1 < K, N not a multiple of 32, batch of 3
    rng = np.random.default_rng(99)
    tiny = rng.random((3, 10, 3), dtype=np.float32)
    arrays["tiny_xyz"] = tiny
    arrays["tiny_k16_self"] = ok.knn_reference(tiny, tiny, 16)
    rag = rng.random((3, 703, 3), dtype=np.float32)
    arrays["ragged_xyz"] = rag
    arrays["ragged_k16_self"] = ok.knn_reference(rag, rag, 16)
    arrays["ragged_k1_prefix"] = ok.knn_reference(rag[:, :351], rag, 1)
    arrays["ragged_k5_cross"] = ok.knn_reference(rag[:, :500], rag[:, 100:], 5)
    for kind, n, seed in LARGE:
        p = make_cloud(kind, n, seed)[None]
        tag = f"{kind}_{n}"
        r16 = ok.knn_reference(p, p, 16)
        sub = p[:, : n // 4]
        r1 = ok.knn_reference(sub, p, 1)
        sums[tag] = dict(kind=kind, n=n, seed=seed, xyz=digest(p), k16_self_idx=digest(r16),
                         k16_self_dist=digest(ok.knn_dists(p, p, r16)), k1_prefix_idx=digest(r1),
                         k1_prefix_dist=digest(ok.knn_dists(sub, p, r1)))
        print(tag, "done")
    np.savez_compressed(os.path.join(HERE, "knn_golden.npz"), **arrays)
    with open(os.path.join(HERE, "knn_checksums.json"), "w") as f:
        json.dump(sums, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
