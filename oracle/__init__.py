"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's PointSegment hot path, used as the *checker* by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
Nothing under ``point_unet_b200/`` may import this package: the product path is CUDA only and
fails loudly when its extension is missing.

Contents
--------
knn_oracle.c / knn.py   C restatement of nanoflann v0x123 kd-tree K-NN + canonical (distance, index)
                        brute force; ctypes bindings; binding of the reference's own compiled C++
                        (``oracle/_ref/libknn_ref.so``, built from /root/reference by ``make ref``).
randla_ref.py           PyTorch-CPU (fp32/fp64) restatement of RandLANet.py / helper_tf_util.py ops.
prepare_ref.py          numpy restatement of the volume -> cloud preparation (dataPreparePancreas.py / dataPrepareBraTS.py /
                        runBraTS.py generator), random draws injected.
(The synthetic Pancreas- / BraTS-shaped cloud generators of SURVEY.md section 8d live in point_unet_b200/synthetic.py.)
"""
