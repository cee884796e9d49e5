"""CPU tests of the TV-L1 oracle (oracle/tvl1.py): the sub-steps that can be pinned in this container
(cv2 grey conversion, dense_flow's 8-bit rule), the committed golden outputs, and the algorithm's own properties."""
import numpy as np
import pytest

from oracle import tvl1
from video_analytics_b200.flow import TVL1Params, synthetic_clip


def test_gray_equals_cv2_on_every_colour():
    cv2 = pytest.importorskip("cv2")
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    rgb = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    assert np.array_equal(tvl1.gray_from_rgb(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))


@pytest.mark.parametrize("shape,dsize", [((240, 320, 3), (340, 256)), ((240, 320), (340, 256)), ((77, 101, 3), (64, 48)),
                                         ((256, 340, 3), (320, 240)), ((31, 17), (200, 300)), ((480, 640, 3), (340, 256))])
def test_frame_resize_equals_cv2(shape, dsize):
    """dense_flow resizes every frame to TSN's new_size before the flow: the restated fixed-point bilinear equals cv2."""
    cv2 = pytest.importorskip("cv2")
    img = np.random.default_rng(sum(shape)).integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(tvl1.resize_linear_u8(img, dsize[1], dsize[0]), cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR))


def test_golden_reproduces(golden):
    g = golden("tvl1_small.npz")
    assert np.array_equal(tvl1.gray_from_rgb(g["clip_a"][0]), g["gray_a_cv2"][0])
    p = tvl1.TVL1Params()
    u1, u2, st = tvl1.tvl1_flow(g["gray_a_cv2"][0], g["gray_a_cv2"][1], p, return_stats=True)
    assert np.array_equal(u1, g["a0_u1"]) and np.array_equal(u2, g["a0_u2"])
    assert st == list(g["a0_iters"])
    assert np.array_equal(tvl1.flow_to_u8(u1, p.bound), g["a0_x"])


def test_flow_to_u8_is_dense_flows_cast():
    u = np.array([-25.0, -20.0, -19.99, 0.0, 0.0784, 0.0785, 10.0, 20.0, 20.01, 3.0588235], np.float32)
    q = tvl1.flow_to_u8(u, 20.0)
    # cvRound(255 * (v + 20) / 40): 0 -> 127.5 -> 128 (half to even), saturation outside [-20, 20]
    assert list(q[:4]) == [0, 0, 0, 128] and q[6] == 191 and q[7] == 255 and q[8] == 255
    assert q[9] == int(np.rint(255.0 * (float(u[9]) + 20.0) / 40.0))


def test_pyramid_sizes_follow_cv_resize_rounding():
    assert tvl1.pyramid_sizes(256, 340, tvl1.TVL1Params()) == [(256, 340), (205, 272), (164, 218), (131, 174), (105, 139)]
    assert tvl1.pyramid_sizes(240, 320, tvl1.TVL1Params()) == [(240, 320), (192, 256), (154, 205), (123, 164), (98, 131)]
    assert tvl1.pyramid_sizes(24, 32, tvl1.TVL1Params()) == [(24, 32), (19, 26)]          # next level would be < 16
    assert TVL1Params().levels(256, 340) == 5 and TVL1Params().levels(24, 32) == 2


def test_recovers_a_known_translation():
    clip = synthetic_clip(2, 64, 96, seed=3, channels=1, velocity=(1.5, -1.0), object_velocity=(1.5, -1.0), noise=0)
    u1, u2 = tvl1.tvl1_flow(clip[0, :, :, 0], clip[1, :, :, 0])
    inner = (slice(12, -12), slice(12, -12))
    assert abs(float(np.median(u1[inner])) - 1.5) < 0.05 and abs(float(np.median(u2[inner])) + 1.0) < 0.05
    qx = tvl1.flow_to_u8(u1, 20.0)
    assert abs(int(np.median(qx[inner])) - round(255 * 21.5 / 40)) <= 1


def test_identical_frames_give_zero_flow_and_stop_early():
    clip = synthetic_clip(1, 48, 64, seed=4, channels=1)
    u1, u2, st = tvl1.tvl1_flow(clip[0, :, :, 0], clip[0, :, :, 0], return_stats=True)
    assert float(np.abs(u1).max()) == 0.0 and float(np.abs(u2).max()) == 0.0
    assert all(n == 2 for n in st)          # first error sum (n = 1) is zero: one more dual update, then out
    assert np.all(tvl1.flow_to_u8(u1, 20.0) == 128)
