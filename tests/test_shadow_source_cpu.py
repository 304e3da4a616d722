"""Shadow-map comparison sampling (shadowed positional lights).  granite_b200/csrc/grb_shadow.cuh
compiled for the CPU (tests/cpp/cuda_host_emul.h) against the oracle's restatement of the Vulkan comparison filter,
bit for bit; and the oracle's cube-edge handling against plain geometry."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_shadow.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{cuda}/include",
           os.path.join(ROOT, "tests", "cpp", "emulate_shadow.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _texel_centre(res, f, i, j):
    """Centre of texel (i, j) of face f on the unit cube (Vulkan's s_c / t_c / m_a table inverted)."""
    a, b = (2 * i + 1 - res) / res, (2 * j + 1 - res) / res
    return {0: (1, -b, -a), 1: (-1, -b, a), 2: (a, 1, b), 3: (a, -1, -b), 4: (a, -b, 1), 5: (-a, -b, -1)}[f]


@pytest.mark.parametrize("res", [4, 7, 16])
def test_oracle_cube_edge_texels_are_the_geometric_neighbours(oracle, res):
    """One step over an edge lands on the texel of the adjacent face that touches the edge texel: the two centres are
    sqrt(2) / res apart on the unit cube (any flipped or transposed mapping is at least a texel further)."""
    lib = oracle.lib()
    lib.orc_shadow_cube_texel.restype = C.c_int
    t = C.c_size_t()
    for f in range(6):
        for k in range(res):
            for (i, j, ei, ej) in [(-1, k, 0, k), (res, k, res - 1, k), (k, -1, k, 0), (k, res, k, res - 1)]:
                assert lib.orc_shadow_cube_texel(res, f, i, j, C.byref(t)) == 1
                nf, rem = divmod(t.value, res * res)
                nj, ni = divmod(rem, res)
                assert nf != f and nf // 2 != f // 2
                d = np.linalg.norm(np.subtract(_texel_centre(res, nf, ni, nj), _texel_centre(res, f, ei, ej)))
                assert abs(d - np.sqrt(2.0) / res) < 1e-6, (f, i, j, nf, ni, nj)
        for (i, j) in [(-1, -1), (res, -1), (-1, res), (res, res)]:
            assert lib.orc_shadow_cube_texel(res, f, i, j, C.byref(t)) == 0
        assert lib.orc_shadow_cube_texel(res, f, 1, 2, C.byref(t)) == 1 and t.value == (f * res + 2) * res + 1


@pytest.mark.parametrize("res", [8, 33, 512])
def test_shadow_2d_source_equals_oracle(emu, oracle, res):
    lib = oracle.lib()
    lib.orc_shadow_sample_2d.restype = C.c_float
    lib.orc_shadow_sample_2d.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    rng = np.random.default_rng(res)
    m = rng.integers(0, 65536, (res, res), dtype=np.uint16)
    n = 4000
    clip = np.empty((n, 4), np.float32)
    clip[:, 3] = rng.uniform(0.2, 30.0, n)
    clip[:, 0] = rng.uniform(-0.1, 1.1, n) * clip[:, 3]
    clip[:, 1] = rng.uniform(-0.1, 1.1, n) * clip[:, 3]
    clip[:, 2] = rng.uniform(-0.1, 1.1, n) * clip[:, 3]
    clip[:8] = [[0, 0, 0.5, 1], [1, 1, 0.5, 1], [0.5, 0.5, 2, 1], [0.5, 0.5, -1, 1], [1, 2, 3, 0], [np.nan, 0, 0, 1], [0.5 / res, 0.5 / res, 0.3, 1],
                [1e30, -1e30, 0.5, 1]]
    out = np.zeros(n, np.float32)
    emu.emu_shadow_2d(_p(m), res, _p(clip), n, _p(out))
    ref = np.array([lib.orc_shadow_sample_2d(_p(m), res, *[float(v) for v in c]) for c in clip], np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert 0.2 < (ref[8:] > 0).mean() < 0.9 and ((ref[8:] > 0) & (ref[8:] < 1)).mean() > 0.2  # partially lit footprints exist


@pytest.mark.parametrize("res", [4, 16, 129])
def test_shadow_cube_source_equals_oracle(emu, oracle, res):
    lib = oracle.lib()
    lib.orc_shadow_sample_cube.restype = C.c_float
    lib.orc_shadow_sample_cube.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    rng = np.random.default_rng(100 + res)
    m = rng.integers(0, 65536, (6, res, res), dtype=np.uint16)
    n = 6000
    d = rng.normal(size=(n, 4)).astype(np.float32)
    d[:, 3] = rng.uniform(-0.1, 1.1, n)
    # a third of the directions hug a cube edge, a sixth a corner (footprints that leave the face)
    k = n // 3
    ax = rng.integers(0, 3, k)
    d[np.arange(k), ax] = np.sign(d[np.arange(k), ax]) * np.abs(d[np.arange(k), (ax + 1) % 3]) * rng.uniform(0.97, 1.03, k).astype(np.float32)
    c = slice(k, k + n // 6)
    mag = np.abs(d[c, 0:1])
    d[c, 0:3] = np.sign(d[c, 0:3]) * mag * rng.uniform(0.97, 1.03, (n // 6, 3)).astype(np.float32)
    d[-4:] = [[1, 1, 1, 0.5], [-1, 1, -1, 0.5], [0, 0, 2, 0.5], [3, -3, 0, 0.5]]
    out = np.zeros(n, np.float32)
    emu.emu_shadow_cube(_p(m), res, _p(d), n, _p(out))
    ref = np.array([lib.orc_shadow_sample_cube(_p(m), res, *[float(v) for v in q]) for q in d], np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert ((ref > 0) & (ref < 1)).mean() > 0.2


def test_cube_sampling_is_continuous_across_edges(oracle):
    """A smooth depth field (depth = a function of direction only) compared against a fixed reference gives the same
    answer, up to one texel's weight, just inside and just outside a face boundary -- no seam."""
    lib = oracle.lib()
    lib.orc_shadow_sample_cube.restype = C.c_float
    lib.orc_shadow_sample_cube.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    res = 32
    m = np.zeros((6, res, res), np.uint16)
    for f in range(6):
        for j in range(res):
            for i in range(res):
                v = np.array(_texel_centre(res, f, i, j), np.float64)
                v /= np.linalg.norm(v)
                m[f, j, i] = int(65535 * (0.5 + 0.45 * np.sin(3 * v[0] + 2 * v[1] - 4 * v[2])))
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(400):
        t = rng.uniform(-0.95, 0.95)
        e = 1e-4
        for (a, b) in [((1.0, 1.0 - e, t), (1.0 - e, 1.0, t)), ((1.0, t, -1.0 + e), (1.0 - e, t, -1.0)), ((t, -1.0, 1.0 - e), (t, -1.0 + e, 1.0))]:
            va = lib.orc_shadow_sample_cube(_p(m), res, *a, 0.5)
            vb = lib.orc_shadow_sample_cube(_p(m), res, *b, 0.5)
            worst = max(worst, abs(va - vb))
    assert worst < 0.02, worst


def test_falloff_helpers_equal_oracle_statements(emu, oracle):
    """spot.h:67-77 / point.h:46-71 around the samplers: the clip transform and the cube reference depth."""
    lib = oracle.lib()
    lib.orc_shadow_sample_2d.restype = C.c_float
    lib.orc_shadow_sample_2d.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    lib.orc_shadow_sample_cube.restype = C.c_float
    lib.orc_shadow_sample_cube.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    rng = np.random.default_rng(9)
    res = 64
    m2 = rng.integers(0, 65536, (res, res), dtype=np.uint16)
    mc = rng.integers(0, 65536, (6, res, res), dtype=np.uint16)
    T = rng.normal(size=16).astype(np.float32)
    T[15] = 5.0
    n = 2000
    pos = rng.normal(size=(n, 3)).astype(np.float32)
    out = np.zeros(n, np.float32)
    emu.emu_spot_shadow_falloff(_p(T), _p(pos), n, _p(m2), res, _p(out))
    f32 = np.float32
    ref = []
    for p in pos:
        c = [(T[r] * p[0] + T[4 + r] * p[1]) + (T[8 + r] * p[2] + T[12 + r] * f32(1.0)) for r in range(4)]
        ref.append(lib.orc_shadow_sample_2d(_p(m2), res, *[float(v) for v in c]))
    assert np.array_equal(out.view(np.uint32), np.array(ref, np.float32).view(np.uint32))
    # point light: near = 0.005 r, far = r  =>  shadow[index][0] = (proj[2].zw, proj[3].zw)
    P = np.array([0.005 / 0.995, -1.0, 0.005 * 10.0 / 0.995, 0.0], np.float32)
    T2 = np.zeros(16, np.float32)
    T2[:4] = P
    full = (rng.normal(size=(n, 3)) * 4).astype(np.float32)
    emu.emu_point_shadow_falloff(_p(T2), _p(full), n, _p(mc), res, _p(out))
    ref = []
    for d in full:
        mz = max(abs(d[0]), abs(d[1]), abs(d[2]))
        rx, ry = T2[2] - T2[0] * mz, T2[3] - T2[1] * mz
        ref.append(lib.orc_shadow_sample_cube(_p(mc), res, float(d[0]), float(d[1]), float(d[2]), float(f32(rx) / f32(ry))))
    assert np.array_equal(out.view(np.uint32), np.array(ref, np.float32).view(np.uint32))


# ---- shadow transforms: host layer == oracle == the reference's own math ----
def _ref_transform(ref, rec, is_point, xy_range):
    out = np.zeros(16, np.float32)
    pos = np.ascontiguousarray(rec["position"], np.float32)
    if is_point:
        ref.ref_point_shadow_transform(_p(pos), C.c_float(float(rec["inv_radius"])), _p(out))
    else:
        d = np.ascontiguousarray(rec["direction"], np.float32)
        ref.ref_spot_shadow_transform(_p(d), _p(pos), C.c_float(float(rec["inv_radius"])), C.c_float(float(xy_range)), _p(out))
    return out


@pytest.mark.parametrize("n,spots", [(16, 0.0), (300, 0.25), (500, 1.0)])
def test_shadow_transforms_host_oracle_reference(oracle, n, spots):
    from granite_b200 import build, synth, viewer
    from tests import common

    build.build_all()
    w, h = 1920, 1080
    v = viewer.Viewer(w, h, cuda_device=-1)
    v.set_camera(synth.perspective_inf(np.pi / 4, w / h, 1 / 16), synth.look_at_view((0, 0, 8), (0, 0, 0)))
    lights = synth.make_lights(n, spot_fraction=spots, aspect=w / h)
    v.set_lights(lights)
    host = v.shadow_transforms()
    cam = common.oracle_camera_from_viewer(oracle, v)
    prep = oracle.prepare_lights(cam, lights)
    orc = oracle.shadow_transforms(prep)
    assert host.shape == orc.shape and prep.n > 0
    assert np.array_equal(host.view(np.uint32), orc.view(np.uint32))
    ref = oracle.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for i in range(prep.n):
        is_point = (int(prep.type_mask[i >> 5]) >> (i & 31)) & 1
        r = _ref_transform(ref, prep.records[i], is_point, oracle.spot_xy_range(prep.outer_cone[i]))
        assert np.array_equal(r.view(np.uint32), orc[i].view(np.uint32)), (i, is_point)
    v.close()


def test_spot_shadow_transform_degenerate_directions(oracle):
    """rotate_vector's special cases (transforms.cpp:128-141): a spot light looking along -Z (identity) and along +Z
    (half a turn), and nearly so."""
    ref = oracle.ref()
    L = oracle.lib()
    for d in [(0, 0, -1), (0, 0, 1), (1e-3, 0, -1), (0, 2e-3, 1), (0.6, 0, 0.8), (1, 0, 0)]:
        rec = oracle.Light()
        rec.direction[:] = list(np.array(d, np.float32) / np.float32(np.linalg.norm(d)))
        rec.position[:] = [1.0, 2.0, 3.0]
        rec.inv_radius = 0.125
        m = np.zeros(16, np.float32)
        L.orc_spot_shadow_transform(C.byref(rec), C.c_float(0.6), _p(m))
        assert np.isfinite(m).all()
        if ref is not None:
            r = np.zeros(16, np.float32)
            ref.ref_spot_shadow_transform(_p(np.array(rec.direction[:], np.float32)), _p(np.array(rec.position[:], np.float32)), C.c_float(0.125), C.c_float(0.6), _p(r))
            assert np.array_equal(r.view(np.uint32), m.view(np.uint32)), d


# ---- SHADOW_MAP_PCF_KERNEL_WIDE (pcf.h:7-80): the 6 x 6 kernel of the spot lights ----
@pytest.mark.parametrize("res", [8, 33, 512])
def test_shadow_2d_wide_source_equals_oracle(emu, oracle, res):
    lib = oracle.lib()
    lib.orc_shadow_sample_2d_wide.restype = C.c_float
    lib.orc_shadow_sample_2d_wide.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    rng = np.random.default_rng(res + 77)
    m = rng.integers(0, 65536, (res, res), dtype=np.uint16)
    n = 3000
    clip = np.empty((n, 4), np.float32)
    clip[:, 3] = rng.uniform(0.2, 30.0, n)
    for c in range(3):
        clip[:, c] = rng.uniform(-0.1, 1.1, n) * clip[:, 3]
    clip[:4] = [[0, 0, 0.5, 1], [1, 1, 0.5, 1], [0.5, 0.5, 2, 1], [0.5 / res, 0.5 / res, 0.3, 1]]
    out = np.zeros(n, np.float32)
    emu.emu_shadow_2d_wide(_p(m), res, _p(clip), n, _p(out))
    ref = np.array([lib.orc_shadow_sample_2d_wide(_p(m), res, *[float(v) for v in c]) for c in clip], np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert np.isfinite(ref).all() and (ref >= 0).all() and (ref <= 1.000001).all() and ((ref > 0.05) & (ref < 0.95)).mean() > 0.3


def test_shadow_2d_wide_is_a_normalised_blur_of_the_comparison(oracle):
    """All texels lit -> 1, none -> 0; across a straight shadow edge the wide kernel ramps over about five texels where
    the sampler's 2 x 2 filter ramps over one."""
    lib = oracle.lib()
    for f in (lib.orc_shadow_sample_2d_wide, lib.orc_shadow_sample_2d):
        f.restype = C.c_float
        f.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
    res = 64
    lit = np.zeros((res, res), np.uint16)
    assert lib.orc_shadow_sample_2d_wide(_p(lit), res, 0.37, 0.61, 0.5, 1.0) == 1.0
    dark = np.full((res, res), 65535, np.uint16)
    assert lib.orc_shadow_sample_2d_wide(_p(dark), res, 0.37, 0.61, 0.5, 1.0) == 0.0
    edge = np.zeros((res, res), np.uint16)
    edge[:, 32:] = 65535
    xs = np.linspace(24, 40, 161) / res
    wide = np.array([lib.orc_shadow_sample_2d_wide(_p(edge), res, float(x), 0.5, 0.5, 1.0) for x in xs])
    narrow = np.array([lib.orc_shadow_sample_2d(_p(edge), res, float(x), 0.5, 0.5, 1.0) for x in xs])
    assert (np.diff(wide) <= 1e-6).all() and wide[0] == 1.0 and wide[-1] == 0.0
    ramp = lambda v: ((v > 0.02) & (v < 0.98)).sum() / 10.0  # texels
    assert 3.5 < ramp(wide) < 6.0 and 0.8 < ramp(narrow) < 1.2, (ramp(wide), ramp(narrow))
