"""CPU tests of the host-side step logic that needs no kernel: the slot layout of the pipelined training step."""
import torch

from point_unet_b200.helper_tool import ConfigBraTS
from point_unet_b200.train import _Slot


def test_slot_layout_is_one_flat_buffer():
    class cfg(ConfigBraTS):
        num_points = 1024
    B, N = 2, 1024
    slot = _Slot(cfg, B, N, 4, torch.int64, torch.device("cpu"))
    st = slot.store
    assert len(st["xyz"]) == cfg.num_layers + 1 and len(st["neigh_idx"]) == len(st["sub_idx"]) == len(st["interp_idx"]) == cfg.num_layers
    n = N
    spans = []
    for i in range(cfg.num_layers):
        n_sub = n // cfg.sub_sampling_ratio[i]
        assert st["xyz"][i].shape == (B, n, 3) and st["xyz"][i].dtype == torch.float32
        assert st["neigh_idx"][i].shape == (B, n, cfg.k_n) and st["neigh_idx"][i].dtype == torch.int32
        assert st["sub_idx"][i].shape == (B, n_sub, cfg.k_n) and st["interp_idx"][i].shape == (B, n, 1)
        (no, np_), (so, sp), (io, ip) = st["inv"][i]
        assert no.numel() == B * n + 1 and np_.numel() == B * n * cfg.k_n       # inverse of neigh_idx: targets = level points
        assert so.numel() == B * n + 1 and sp.numel() == B * n_sub * cfg.k_n    # inverse of sub_idx (pool): same targets
        assert io.numel() == B * n_sub + 1 and ip.numel() == B * n              # inverse of interp_idx: targets = sub-cloud
        n = n_sub
    assert st["xyz"][-1].shape == (B, n, 3)
    assert slot.features.shape == (B, N, 4) and slot.labels.shape == (B, N) and slot.labels.dtype == torch.int64
    base = slot.flat.data_ptr()
    for t in slot.t.values():
        off = t.data_ptr() - base
        assert off % 256 == 0 and t.is_contiguous()
        spans.append((off, off + t.numel() * t.element_size()))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= slot.flat.numel()
    # handing a slot over = one copy of the flat buffer
    other = _Slot(cfg, B, N, 4, torch.int64, torch.device("cpu"))
    for k, t in slot.t.items():
        t.copy_(torch.randint(0, 100, t.shape).to(t.dtype))
    other.flat.copy_(slot.flat)
    assert all(torch.equal(other.t[k], slot.t[k]) for k in slot.t)
    pyr = other.pyramid()
    assert len(pyr["xyz"]) == cfg.num_layers and pyr["xyz"][0].data_ptr() == other.store["xyz"][0].data_ptr()
