"""parameters.py carries the reference's 43 constants unchanged (golden dumped from Sheet03/parameters.py)."""
from video_analytics_b200 import parameters as P


def test_constants_match_reference(golden):
    ref = golden("reference_parameters.json")
    assert len(ref) == 43
    for name, value in ref.items():
        assert hasattr(P, name), name
        assert getattr(P, name) == value, name


def test_star_import_surface():
    ns = {}
    exec("from video_analytics_b200.parameters import *", ns)
    for name in ("CROP_SIZE_TF", "NORM_MEANS_TF", "VIDEO_DESCRIPTOR_DIM", "N_TEST_SNIPPETS", "N_TEST_CROPS"):
        assert name in ns
    assert ns["N_TEST_SNIPPETS"] * ns["N_TEST_CROPS"] == 250
