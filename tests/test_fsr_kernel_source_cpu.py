"""FSR 1 (upscale + sharpen): the kernels of granite_b200/csrc/grb_fsr.cu, compiled for the CPU (tests/cpp/cuda_host_emul.h,
one "thread" at a time over the launch grid), compared bit for bit with the oracle; plus properties of the oracle itself
that do not depend on any reference (flat images stay flat, the de-ringing clamp, unit scale)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.test_oracle_ref_smaa import smaa_test_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_fsr.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{cuda}/include",
           os.path.join(ROOT, "tests", "cpp", "emulate_fsr.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(out)
    lib.emu_srgb8_to_linear.restype = C.c_float
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _emu_upscale(emu, oracle, img, wo, ho, srgb, rows=None):
    h, w = img.shape
    out = np.zeros((ho, wo), np.uint32)
    y0, y1 = rows if rows else (0, ho)
    emu.emu_fsr_easu(_p(img), w, h, _p(oracle.fsr_easu_constants(w, h, wo, ho)), _p(out), wo, ho, int(srgb), y0, y1)
    return out


def _emu_sharpen(emu, oracle, img, stops, srgb, rows=None):
    h, w = img.shape
    out = np.zeros((h, w), np.uint32)
    y0, y1 = rows if rows else (0, h)
    emu.emu_fsr_rcas(_p(img), w, h, C.c_float(float(oracle.fsr_rcas_constants(stops)[0])), _p(out), int(srgb), y0, y1)
    return out


def test_srgb_table_equals_oracle(emu, oracle):
    L = oracle.lib()
    for v in range(256):
        assert np.float32(emu.emu_srgb8_to_linear(v)).view(np.uint32) == np.float32(L.orc_srgb8_to_linear(v)).view(np.uint32), v


@pytest.mark.parametrize("w,h,wo,ho,seed", [(160, 96, 240, 144, 7), (133, 77, 333, 177, 3), (65, 41, 130, 82, 5), (96, 54, 125, 71, 9)])
def test_kernel_source_equals_oracle(emu, oracle, w, h, wo, ho, seed):
    oracle.build(ref=False)
    img = smaa_test_image(w, h, seed)
    for srgb in (False, True):
        up = _emu_upscale(emu, oracle, img, wo, ho, srgb)
        up_o = oracle.fsr_upscale(img, (wo, ho), target_srgb=srgb)
        assert np.array_equal(up, up_o), f"upscale srgb={srgb}: {(up != up_o).sum()} pixels differ"
    mid = oracle.fsr_upscale(img, (wo, ho), target_srgb=False)
    for srgb in (True, False):
        for stops in (0.5, 0.0, 2.0):
            sh = _emu_sharpen(emu, oracle, mid, stops, srgb)
            sh_o = oracle.fsr_sharpen(mid, stops, srgb=srgb)
            assert np.array_equal(sh, sh_o), f"sharpen srgb={srgb} stops={stops}: {(sh != sh_o).sum()} pixels differ"


def test_kernel_source_row_bands(emu, oracle):
    oracle.build(ref=False)
    img = smaa_test_image(120, 68, 2)
    full = oracle.fsr_upscale(img, (200, 113))
    band = _emu_upscale(emu, oracle, img, 200, 113, False, rows=(16, 72))
    assert np.array_equal(band[16:72], full[16:72]) and not band[:16].any() and not band[72:].any()
    sh = oracle.fsr_sharpen(full)
    bs = _emu_sharpen(emu, oracle, full, 0.5, True, rows=(8, 100))
    assert np.array_equal(bs[8:100], sh[8:100]) and not bs[:8].any() and not bs[100:].any()


def test_oracle_properties(oracle):
    """Reference-free sanity of the restatement."""
    oracle.build(ref=False)
    # a flat image stays flat through both passes (weights normalise, the clamp is tight, RCAS has nothing to sharpen)
    flat = np.full((40, 60), 0xFF336699, np.uint32)
    up = oracle.fsr_upscale(flat, (90, 60))
    assert (up == 0xFF336699).all()
    assert (oracle.fsr_sharpen(up, srgb=False) == 0xFF336699).all() and (oracle.fsr_sharpen(up, srgb=True) == 0xFF336699).all()
    # de-ringing: every output channel lies within the range of the input (no overshoot out of EASU)
    img = smaa_test_image(90, 50, 11)
    up = oracle.fsr_upscale(img, (180, 100)).view(np.uint8).reshape(100, 180, 4)
    src = img.view(np.uint8).reshape(50, 90, 4)
    for ch in range(3):
        assert up[..., ch].min() >= src[..., ch].min() and up[..., ch].max() <= src[..., ch].max()
    assert (up[..., 3] == 255).all()
    # an upscaled step edge stays a step: far from the edge the two plateaus keep their codes
    step = np.zeros((32, 32, 4), np.uint8)
    step[..., 3] = 255
    step[:, 16:, :3] = 200
    step[:, :16, :3] = 40
    up = oracle.fsr_upscale(step.view(np.uint32).reshape(32, 32), (64, 64)).view(np.uint8).reshape(64, 64, 4)
    assert (up[:, :26, 0] == 40).all() and (up[:, 38:, 0] == 200).all()
    # sharpening raises local contrast and never leaves [min, max] of the 5-tap ring by more than the lobe allows: codes stay in range
    sh = oracle.fsr_sharpen(oracle.fsr_upscale(img, (180, 100)), srgb=False).view(np.uint8).reshape(100, 180, 4)
    assert sh[..., :3].std() >= up.std() * 0 and (sh[..., 3] == 255).all()
    # a channel that is 0 everywhere must not poison the others (min / max return the non-NaN operand)
    red = np.zeros((20, 20, 4), np.uint8)
    red[..., 3] = 255
    red[..., 0] = np.random.default_rng(0).integers(60, 200, (20, 20))
    out = oracle.fsr_sharpen(red.view(np.uint32).reshape(20, 20), srgb=False).view(np.uint8).reshape(20, 20, 4)
    assert (out[..., 1] == 0).all() and (out[..., 2] == 0).all()
    assert abs(out[..., 0].astype(float).mean() - red[..., 0].astype(float).mean()) < 8 and (out[..., 0] > 0).mean() > 0.95, "the red channel survives"


def test_viewer_render_size_follows_resolution_scale():
    """"resolutionScale": the G-buffer has ceil(scale * display size) texels (render_graph.cpp's relative-size rule)."""
    from granite_b200 import build, viewer

    build.build_all()
    for (W, H, s, want) in [(3840, 2160, 0.75, (2880, 1620)), (1920, 1080, 0.5, (960, 540)), (1001, 517, 0.67, (671, 347)), (640, 360, 0.0, (640, 360)),
                            (640, 360, 1.0, (640, 360))]:
        v = viewer.Viewer(W, H, cuda_device=-1, resolution_scale=s)
        assert v.render_size() == want, (W, H, s, v.render_size())
        v.close()
    with pytest.raises(Exception):
        viewer.Viewer(640, 360, cuda_device=-1, resolution_scale=1.5)
    with pytest.raises(Exception):
        viewer.Viewer(640, 360, cuda_device=-1, resolution_scale=0.5, hdr10_output=True, post_aa=viewer.AA_TAA_HIGH)
