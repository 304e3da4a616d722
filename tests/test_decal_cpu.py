"""Volumetric-decal binning (clusterer.cpp:1348-1461, clusterer_bindless_binning_decal.comp): the oracle
pinned bit for bit to the reference's own shader run on the CPU (its SUBGROUPS=0 path), the kernels of
granite_b200/csrc/grb_decal.cu compiled for the CPU bit for bit against the oracle, and the bitmask against plain geometry."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_decals(n, seed=3, aspect=16.0 / 9.0):
    """World transforms (mat_affine rows) of n decal boxes: random rotation, 0.5 .. 6 m edge lengths, spread through the
    view frustum of the synthetic camera (eye (0, 0, 8) looking down -Z), some behind the camera and some straddling it."""
    rng = np.random.default_rng(seed + n)
    rows = np.zeros((n, 12), np.float32)
    for i in range(n):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        S = np.diag(rng.uniform(0.5, 6.0, 3))
        depth = rng.uniform(1.0, 120.0) if (i + 1) % 9 else rng.uniform(-10.0, 1.0)  # every 9th: behind / across the camera plane
        half = depth * np.tan(np.pi / 8)
        pos = np.array([rng.uniform(-1.2, 1.2) * half * aspect, rng.uniform(-1.2, 1.2) * half, 8.0 - depth])
        M = np.concatenate([R @ S, pos[:, None]], axis=1)
        rows[i] = M.astype(np.float32).reshape(-1)
    return rows


def _camera(oracle, aspect=16.0 / 9.0):
    cam, _, _ = common.build_lights_case(oracle, aspect, 16)
    return cam


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_decal.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{cuda}/include",
           os.path.join(ROOT, "tests", "cpp", "emulate_decal.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(out)


@pytest.mark.parametrize("n,res", [(1, (16, 8)), (40, (128, 64)), (300, (64, 32))])
def test_oracle_decal_binning_equals_reference_shader(oracle, n, res):
    oracle.build()
    k = oracle.ref_kernels()
    if k is None or 5 not in k:
        pytest.skip("oracle/_ref shaders not built (no /root/reference on this machine)")
    cam = _camera(oracle)
    mvps = oracle.decal_mvps(cam, make_decals(n))
    mine, ref = oracle.decal_binning(res, mvps), oracle.ref_decal_binning(res, mvps)
    assert np.array_equal(mine, ref)
    filled = np.unpackbits(mine.view(np.uint8)).mean()
    assert filled < 0.6 and (n == 1 or filled > 0.001), filled


@pytest.mark.parametrize("n,res", [(1, (16, 8)), (33, (128, 64)), (300, (64, 32)), (4096, (32, 16))])
def test_kernel_source_equals_oracle(emu, oracle, n, res):
    cam = _camera(oracle)
    mvps = oracle.decal_mvps(cam, make_decals(n))
    rx, ry = res
    boxes = np.zeros((n, 4), np.float32)
    got = np.zeros((ry, rx, (n + 31) // 32), np.uint32)
    emu.emu_decal_binning(mvps.ctypes.data_as(C.c_void_p), n, rx, ry, C.c_float(np.float32(1.0 / rx)), C.c_float(np.float32(1.0 / ry)),
                          boxes.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, oracle.decal_binning(res, mvps))
    bb = np.zeros(4, np.float32)
    for i in range(0, n, max(n // 50, 1)):
        oracle.lib().orc_decal_screen_bb(mvps[i].ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p))
        assert np.array_equal(bb.view(np.uint32), boxes[i].view(np.uint32)), i


def test_decal_bitmask_is_conservative_and_tight(oracle):
    """Every tile a decal's projected corners fall into has the decal's bit; tiles more than the box away do not."""
    cam = _camera(oracle)
    rows = make_decals(60, seed=11)
    mvps = oracle.decal_mvps(cam, rows)
    rx, ry = 128, 64
    bm = oracle.decal_binning((rx, ry), mvps)
    for i in range(60):
        m = mvps[i].reshape(4, 4).T.astype(np.float64)  # column-major -> matrix
        corners = np.array([[sx, sy, sz, 1.0] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)]) @ m.T
        bit = (bm[:, :, i >> 5] >> np.uint32(i & 31)) & 1
        if (corners[:, 3] <= 0).all():
            assert not bit.any(), i
            continue
        if (corners[:, 3] <= 0).any():
            assert bit.all(), i  # straddles the camera plane: the whole screen
            continue
        ndc = corners[:, :2] / corners[:, 3:4]
        for (u, v) in ndc:
            if -1 <= u < 1 and -1 <= v < 1:
                assert bit[int((v + 1) / 2 * ry), int((u + 1) / 2 * rx)] == 1, i
        lo, hi = ndc.min(0), ndc.max(0)
        ys, xs = np.nonzero(bit)
        if len(xs):
            assert (2 * (xs + 1) / rx - 1 > lo[0] - 1e-5).all() and (2 * xs / rx - 1 < hi[0] + 1e-5).all(), i
            assert (2 * (ys + 1) / ry - 1 > lo[1] - 1e-5).all() and (2 * ys / ry - 1 < hi[1] + 1e-5).all(), i


def test_decal_z_range_equals_a_numpy_scan(oracle):
    cam = _camera(oracle)
    rows = make_decals(80, seed=5)
    zr = oracle.decal_z_ranges(cam, rows)
    pos, front = np.array(list(cam.camera_position), np.float64), np.array(list(cam.camera_front), np.float64)
    for i in range(80):
        M = rows[i].reshape(3, 4).astype(np.float64)
        c = np.array([[sx, sy, sz, 1.0] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)]) @ M.T
        z = (c - pos) @ front
        assert abs(zr[i, 0] - z.min()) < 1e-3 and abs(zr[i, 1] - z.max()) < 1e-3


def test_host_decal_prep_equals_oracle(oracle):
    """LightClusterer's decal preparation (frustum cull, depth sort, view_projection * world, Z-slice ranges) against the
    oracle's restatement of clusterer.cpp:1348-1410, bit for bit."""
    from granite_b200 import build, synth, viewer

    build.build_all()
    w, h = 1920, 1080
    v = viewer.Viewer(w, h, cuda_device=-1)
    v.set_camera(synth.perspective_inf(np.pi / 4, w / h, 1 / 16), synth.look_at_view((0, 0, 8), (0, 0, 0)))
    rows = make_decals(200, seed=21)
    v.set_decals(rows)
    mvps, zr = v.decal_prep()
    cam = common.oracle_camera_from_viewer(oracle, v)
    assert 20 < len(mvps) < 200, "some decals are culled, most are visible"
    # identify which decal each visible entry is through its mvp, then compare the whole record
    all_mvps = oracle.decal_mvps(cam, rows)
    index = {all_mvps[i].tobytes(): i for i in range(len(rows))}
    ids = [index[m.tobytes()] for m in mvps]  # KeyError = the host's mat4 product differs from the oracle's
    assert len(set(ids)) == len(ids)
    # front to back by the view depth of the box centre
    front, depth = np.array(list(cam.camera_front), np.float64), []
    for i in ids:
        M = rows[i].reshape(3, 4).astype(np.float64)
        c = np.array([[sx, sy, sz, 1.0] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)]) @ M.T
        depth.append(((c.min(0) + c.max(0)) / 2) @ front)
    assert (np.diff(depth) > -1e-3).all()
    # Z-slice ranges: clusterer.cpp:1371-1389 through compute_uint_range (slice extent 0.5 m, 4096 slices)
    lohi = oracle.decal_z_ranges(cam, rows[ids])
    for k in range(len(ids)):
        lo, hi = np.float32(lohi[k, 0]) / np.float32(0.5), np.float32(lohi[k, 1]) / np.float32(0.5)
        want = (0xFFFFFFFF, 0) if hi < 0 else (int(max(lo, 0.0)), min(int(hi), 4095))
        assert tuple(int(x) for x in zr[k]) == want, (k, zr[k], want)
    v.close()
