"""FSR 1: the oracle (oracle/oracle_fsr.c) pinned to the REFERENCE's own upscale.frag / sharpen.frag
(assets/shaders/post/ffx-fsr over ffx_fsr1.h and ffx_a.h -> SPIR-V with the reference's vendored glslang -> C++ with its
vendored spirv-cross -> executed per pixel on the CPU, oracle/ref_post_shim.cpp KERNEL 24 / 25 / 26), and its constant
blocks pinned to the reference's host math.  8-bit codes are compared exactly."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.test_oracle_ref_smaa import smaa_test_image

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_or_skip(oracle):
    oracle.build()
    k = oracle.ref_post_kernels()
    if k is None or 24 not in k:
        pytest.skip("oracle/_ref post shaders are not available on this machine (no /root/reference)")


@pytest.mark.parametrize("w,h,wo,ho,seed", [(160, 96, 240, 144, 7), (133, 77, 333, 177, 3), (96, 54, 125, 71, 9)])
def test_oracle_fsr_equals_reference_shaders(oracle, w, h, wo, ho, seed):
    _ref_or_skip(oracle)
    img = smaa_test_image(w, h, seed)
    for srgb in (False, True):
        mine, ref = oracle.fsr_upscale(img, (wo, ho), target_srgb=srgb), oracle.ref_fsr_upscale(img, (wo, ho), target_srgb=srgb)
        assert np.array_equal(mine, ref), f"upscale srgb={srgb}: {(mine != ref).sum()} of {mine.size} pixels differ"
    mid = oracle.fsr_upscale(img, (wo, ho))
    for srgb in (True, False):
        for stops in (0.5, 0.0):
            mine, ref = oracle.fsr_sharpen(mid, stops, srgb=srgb), oracle.ref_fsr_sharpen(mid, stops, srgb=srgb)
            assert np.array_equal(mine, ref), f"sharpen srgb={srgb} stops={stops}: {(mine != ref).sum()} pixels differ"
            assert (mine != mid).mean() > 0.05, "the pass must actually sharpen"


def test_oracle_fsr_zero_channel_ring_matches_reference(oracle):
    """A channel that is 0 over a whole 5-tap ring (hitMin = 0 * inf): both sides use a GPU's min / max."""
    _ref_or_skip(oracle)
    rng = np.random.default_rng(4)
    img = np.zeros((24, 40, 4), np.uint8)
    img[..., 0] = rng.integers(0, 256, (24, 40))
    img[:, 20:, 1] = rng.integers(0, 256, (24, 20))
    img[..., 3] = 255
    u = np.ascontiguousarray(img).view(np.uint32).reshape(24, 40)
    for srgb in (True, False):
        assert np.array_equal(oracle.fsr_sharpen(u, 0.5, srgb=srgb), oracle.ref_fsr_sharpen(u, 0.5, srgb=srgb))


def test_oracle_reproduces_reference_fsr_fixture(oracle):
    """Runs everywhere: the reference shaders' images come from the committed fixture."""
    f = np.load(os.path.join(GOLDEN, "reffsr_160x96_to_240x144.npz"))
    img = smaa_test_image(160, 96, 7)
    assert np.array_equal(img, f["color"]), "test image generator changed: regenerate the fixture"
    up = oracle.fsr_upscale(img, (240, 144))
    assert np.array_equal(up, f["upscaled_unorm"]) and np.array_equal(oracle.fsr_upscale(img, (240, 144), target_srgb=True), f["upscaled_srgb"])
    assert np.array_equal(oracle.fsr_sharpen(up, 0.5, srgb=True), f["sharpened_srgb"]) and np.array_equal(oracle.fsr_sharpen(up, 0.5, srgb=False), f["sharpened_unorm"])


def test_fsr_constants_host_oracle(oracle):
    """FsrEasuCon / FsrRcasCon (aa.cpp:33-73): the C ABI's constants equal the oracle's; 2^-0.5 and its half."""
    from granite_b200 import build, capi

    build.build_all()
    lib = C.CDLL(capi.LIB_PATH)
    for (w, h, wo, ho) in [(1280, 720, 1920, 1080), (1477, 831, 3840, 2160), (160, 96, 240, 144), (333, 177, 334, 178)]:
        got = np.zeros(16, np.float32)
        assert lib.grb_fsr_easu_constants(w, h, wo, ho, got.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(got.view(np.uint32), oracle.fsr_easu_constants(w, h, wo, ho).view(np.uint32))
    con = oracle.fsr_rcas_constants(0.5)
    assert con[0] == np.float32(2.0 ** -0.5) and con.view(np.uint32)[1] == 0x39A839A8 and con[2] == 0 and con[3] == 0
