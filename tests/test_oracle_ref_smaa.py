"""The oracle's SMAA passes (oracle_smaa.c: luma edge detection, blending-weight calculation with diagonal and corner
detection, neighbourhood blending; presets Low .. Ultra) pinned to the REFERENCE's own shaders
(assets/shaders/post/smaa_*.frag + SMAA.hlsl -> SPIR-V -> C++ on the CPU, oracle/ref_post_shim.cpp KERNEL 150-173) and
to its own lookup textures (assets/textures/smaa/{area,search}.gtx).  Every stored value bit for bit.

Live tests need /root/reference; the fixture test replays vectors those executables wrote
(tests/golden/refsmaa_160x96.npz, made by tests/golden/make_ref_smaa_golden.py) and runs everywhere."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def smaa_test_image(w, h, seed=0):
    """Flat regions separated by horizontal, vertical, shallow, 45-degree and curved edges, some thin features and
    a noisy patch (crossing edges, corners), as 8-bit sRGB-encoded RGBA."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 3), np.float32) + 0.15
    img[(xx * 0.35 + yy) > h * 0.8] = (0.9, 0.7, 0.2)
    img[((xx - w * 0.3) ** 2 + (yy - h * 0.4) ** 2) < (h * 0.22) ** 2] = (0.2, 0.6, 0.9)
    img[(yy > h * 0.1) & (yy < h * 0.2) & (xx > w * 0.55) & (xx < w * 0.9)] = (0.8, 0.8, 0.8)
    img[(xx - yy * 1.0) > w * 0.75] = (0.1, 0.9, 0.3)
    img[(yy % 7 == 3) & (xx > w * 0.05) & (xx < w * 0.25) & (yy > h * 0.7)] = (1.0, 1.0, 1.0)   # one-pixel lines
    img[((xx + yy) % 9 == 0) & (xx > w * 0.6) & (yy > h * 0.55) & (yy < h * 0.75)] = (0.0, 0.0, 0.0)  # thin diagonals
    patch = (xx > w * 0.4) & (xx < w * 0.55) & (yy > h * 0.05) & (yy < h * 0.3)
    img[patch] = rng.random((int(patch.sum()), 3)).astype(np.float32)
    img += rng.normal(0, 0.004, img.shape).astype(np.float32)
    rgba = np.concatenate([np.clip(img, 0, 1), rng.random((h, w, 1)).astype(np.float32)], -1)
    return np.ascontiguousarray((rgba * 255 + 0.5).astype(np.uint8)).view(np.uint32).reshape(h, w)


def _ref_or_skip(oracle):
    oracle.build()
    if oracle.ref_post_kernels() is None or not os.path.isdir(oracle.SMAA_LUT_DIR):
        pytest.skip("oracle/_ref post shaders / the reference's SMAA textures are not available on this machine")


@pytest.mark.parametrize("quality", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,seed", [(160, 96, 0), (333, 177, 1)])
def test_oracle_smaa_equals_reference_shaders(oracle, quality, w, h, seed):
    _ref_or_skip(oracle)
    area, search = oracle.smaa_luts()
    assert area.shape == (560, 160, 2) and search.shape == (16, 64, 1)
    img = smaa_test_image(w, h, seed)
    e_ref = oracle.ref_smaa_edge(img, quality)
    assert np.array_equal(oracle.smaa_edge(img, quality), e_ref)
    assert (e_ref > 0).sum() > 200  # the image does have edges
    w_ref = oracle.ref_smaa_weights(e_ref, area, search, quality)
    assert np.array_equal(oracle.smaa_weights(e_ref, area, search, quality), w_ref)
    assert (w_ref != 0).sum() > 200
    b_ref = oracle.ref_smaa_blend(img, w_ref, quality)
    assert np.array_equal(oracle.smaa_blend(img, w_ref), b_ref)
    # rows: a band computes exactly its rows of the whole image
    band = (8, h - 16)
    assert np.array_equal(oracle.smaa_weights(e_ref, area, search, quality, rows=band)[band[0]:band[1]], w_ref[band[0]:band[1]])


def test_smaa_leaves_flat_regions_alone(oracle):
    """No edge, no weight: the blend pass returns the colour it fetched (decode to linear, encode on store: exact)."""
    oracle.build()
    flat = np.full((32, 48), 0xFF7F4020, np.uint32)
    e = oracle.smaa_edge(flat, 3)
    assert not e.any()
    area = np.zeros((560, 160, 2), np.uint8)
    search = np.zeros((16, 64, 1), np.uint8)
    wgt = oracle.smaa_weights(e, area, search, 3)
    assert not wgt.any()
    assert np.array_equal(oracle.smaa_blend(flat, wgt), flat)


def test_oracle_reproduces_reference_smaa_fixture(oracle):
    oracle.build()
    f = np.load(os.path.join(GOLDEN, "refsmaa_160x96.npz"))
    for q in range(4):
        e = oracle.smaa_edge(f["color"], q)
        assert np.array_equal(e, f[f"q{q}_edges"])
        wgt = oracle.smaa_weights(f[f"q{q}_edges"], f["area"], f["search"], q)
        assert np.array_equal(wgt, f[f"q{q}_weights"])
        assert np.array_equal(oracle.smaa_blend(f["color"], f[f"q{q}_weights"]), f[f"q{q}_out"])
