"""Host-side logic around the TV-L1 flow producer and the round-2 plumbing (no GPU): level planning against the oracle,
the reference's flow-tree naming, the fused front end's support predicate, the NCCL CTA reservation."""
import os

import numpy as np
import pytest
import torch

from oracle import tvl1 as otv
from video_analytics_b200 import flow


@pytest.mark.parametrize("hw", [(256, 340), (240, 320), (48, 64), (33, 41), (24, 32), (16, 16), (480, 640), (100, 77)])
def test_level_count_matches_the_oracle_pyramid(hw):
    for nscales, step in ((5, 0.8), (3, 0.7), (8, 0.5), (1, 0.8)):
        p = flow.TVL1Params(nscales=nscales, scale_step=step)
        op = otv.TVL1Params(nscales=nscales, scale_step=step)
        assert p.levels(*hw) == len(otv.pyramid_sizes(hw[0], hw[1], op))


def test_params_struct_and_output_shape():
    import ctypes as C
    p = flow.TVL1Params(new_size=(340, 256), bound=15.0)
    c = p._c()
    assert C.sizeof(c) == 6 * 8 + 6 * 4                      # va_tvl1_params in include/va_b200.h: 6 doubles + 6 ints
    assert (c.resize_w, c.resize_h, c.bound, c.nscales) == (340, 256, 15.0, 5)
    assert p.out_shape(240, 320) == (256, 340) and flow.TVL1Params().out_shape(240, 320) == (240, 320)


def test_synthetic_clip_is_deterministic_and_moves():
    a = flow.synthetic_clip(3, 32, 48, seed=5)
    b = flow.synthetic_clip(3, 32, 48, seed=5)
    assert a.shape == (3, 32, 48, 3) and a.dtype == np.uint8 and np.array_equal(a, b)
    assert not np.array_equal(a[0], a[1]) and not np.array_equal(a, flow.synthetic_clip(3, 32, 48, seed=6))


def test_flow_tree_has_the_names_the_reference_lists(tmp_path):
    """TemporalDataset counts len(os.listdir(dir)) / 2 images and opens flow_x_%04d / flow_y_%04d from 1
    (temporalModel.py:76-81, parameters.py:38-39)."""
    cv2 = pytest.importorskip("cv2")
    from video_analytics_b200 import parameters as P
    fx = torch.arange(3 * 20 * 24, dtype=torch.int64).reshape(3, 20, 24).remainder(251).to(torch.uint8)
    fy = 255 - fx
    flow.write_flow_tree(str(tmp_path / "v"), fx, fy, quality=100)
    names = sorted(os.listdir(tmp_path / "v"))
    assert len(names) // 2 == 3
    px, py = P.FLOW_X_PREFIX if hasattr(P, "FLOW_X_PREFIX") else "flow_x_", P.FLOW_Y_PREFIX if hasattr(P, "FLOW_Y_PREFIX") else "flow_y_"
    assert names == [f"{px}{k:04d}.jpg" for k in (1, 2, 3)] + [f"{py}{k:04d}.jpg" for k in (1, 2, 3)]
    back = cv2.imread(str(tmp_path / "v" / f"{px}0002.jpg"), cv2.IMREAD_GRAYSCALE)
    assert back.shape == (20, 24) and np.abs(back.astype(int) - fx[1].numpy().astype(int)).max() <= 2


def test_flow_calls_refuse_cpu_tensors():
    from video_analytics_b200._lib import VAError
    with pytest.raises(VAError):
        flow.tvl1(torch.zeros(64, dtype=torch.uint8), (4, 4, 1), torch.zeros((1, 4), dtype=torch.int32), torch.zeros(64, dtype=torch.uint8))


def test_fused_front_end_support_predicate():
    from video_analytics_b200 import ops
    ok = ops.StreamNet.forward_store_supported
    rgb = torch.zeros(240 * 320 * 3 + 64, dtype=torch.uint8)
    base = rgb[(-rgb.data_ptr()) % 16:]                       # 16-byte aligned view
    assert ok(base, (240, 320, 3), 1) and ok(base, (256, 340, 1), 20) and ok(base, (256, 340, 1), 23)
    assert not ok(base, (256, 340, 1), 24)                   # strips of 24 planes do not fit beside the operand ring
    assert not ok(base, (240, 320, 3), 2) and not ok(base, (200, 320, 3), 1)
    assert not ok(base, (225, 227, 1), 20)                   # 225*227 bytes per image: not a multiple of 16
    assert not ok(base[1:], (240, 320, 3), 1)                # unaligned store


def test_nccl_cta_reservation_is_opt_in(monkeypatch):
    from video_analytics_b200 import distributed as D
    monkeypatch.delenv("NCCL_MAX_CTAS", raising=False)
    monkeypatch.delenv("VA_ALLREDUCE_SMS", raising=False)
    D.reserve_nccl_ctas()
    assert os.environ["NCCL_MAX_CTAS"] == "8"
    monkeypatch.setenv("NCCL_MAX_CTAS", "24")
    D.reserve_nccl_ctas()                                    # a user's own setting wins
    assert os.environ["NCCL_MAX_CTAS"] == "24"
