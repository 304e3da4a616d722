"""The SMAA kernels of granite_b200/csrc/grb_smaa.cu, compiled for the CPU (tests/cpp/cuda_host_emul.h: the *_rn
intrinsics as single IEEE operations, one "thread" at a time over the launch grid) and compared bit for bit with the
oracle and with the reference-shader fixture.  This checks the kernel SOURCE -- indexing, control flow, arithmetic
order, partial blocks at odd image sizes -- on machines without a GPU; the GPU run of the same kernels is
tests/test_zz_gpu_smaa.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.test_oracle_ref_smaa import smaa_test_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_smaa.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{cuda}/include",
           os.path.join(ROOT, "tests", "cpp", "emulate_smaa.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run(emu, img, area, search, q, srgb=1):
    h, w = img.shape
    e = np.zeros((h, w, 2), np.uint8)
    emu.emu_smaa_edge(_p(img), w, h, q, _p(e), 0, h)
    wg = np.zeros((h, w), np.uint32)
    emu.emu_smaa_weights(_p(e), w, h, _p(area), _p(search), q, _p(wg), 0, h)
    out = np.zeros((h, w), np.uint32)
    emu.emu_smaa_blend(_p(img), _p(wg), w, h, srgb, _p(out), 0, h)
    return e, wg, out


@pytest.mark.parametrize("w,h,seed", [(160, 96, 7), (333, 177, 3), (65, 41, 5)])
def test_kernel_source_equals_oracle(emu, oracle, w, h, seed):
    oracle.build(ref=False)
    f = np.load(os.path.join(GOLDEN, "refsmaa_160x96.npz"))
    area, search = np.ascontiguousarray(f["area"]), np.ascontiguousarray(f["search"])
    img = smaa_test_image(w, h, seed)
    for q in range(4):
        e, wg, out = _run(emu, img, area, search, q)
        e_o = oracle.smaa_edge(img, q)
        w_o = oracle.smaa_weights(e_o, area, search, q)
        assert np.array_equal(e, e_o), f"edges q{q}"
        assert np.array_equal(wg, w_o), f"weights q{q}"
        assert np.array_equal(out, oracle.smaa_blend(img, w_o)), f"blend q{q}"


def test_kernel_source_reproduces_reference_shader_fixture(emu):
    f = np.load(os.path.join(GOLDEN, "refsmaa_160x96.npz"))
    area, search = np.ascontiguousarray(f["area"]), np.ascontiguousarray(f["search"])
    for q in range(4):
        e, wg, out = _run(emu, np.ascontiguousarray(f["color"]), area, search, q)
        assert np.array_equal(e, f[f"q{q}_edges"]) and np.array_equal(wg, f[f"q{q}_weights"]) and np.array_equal(out, f[f"q{q}_out"])


def test_kernel_source_rows_and_unorm_target(emu, oracle):
    """A band writes only its rows; a UNORM target stores the blended colour without the sRGB round trip."""
    oracle.build(ref=False)
    f = np.load(os.path.join(GOLDEN, "refsmaa_160x96.npz"))
    img = np.ascontiguousarray(f["color"])
    h, w = img.shape
    wg = np.ascontiguousarray(f["q3_weights"])
    out = np.zeros((h, w), np.uint32)
    emu.emu_smaa_blend(_p(img), _p(wg), w, h, 1, _p(out), 16, 72)
    assert np.array_equal(out[16:72], f["q3_out"][16:72]) and not out[:16].any() and not out[72:].any()
    lin = np.zeros((h, w), np.uint32)
    emu.emu_smaa_blend(_p(img), _p(wg), w, h, 0, _p(lin), 0, h)
    untouched = wg == 0
    # where no neighbour contributes a weight either, the colour passes through
    a = wg.view(np.uint8).reshape(h, w, 4)
    right = np.zeros((h, w), bool); right[:, :-1] = a[:, 1:, 3] > 0
    below = np.zeros((h, w), bool); below[:-1] = a[1:, :, 1] > 0
    keep = untouched & ~right & ~below
    assert keep.sum() > 0.8 * h * w and np.array_equal(lin[keep], img[keep])
