"""CPU-only checks of the oracle itself: pinned against the reference's own math/ where that
compiles (oracle/_ref), against numpy for the storage formats, and against size-independent
properties of the algorithms it restates."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from granite_b200 import synth
from tests import common


# --------------------------------------------------------------------------------------
# host math pinned bit-for-bit against the reference's muglm (oracle/_ref)
# --------------------------------------------------------------------------------------
def _ref_or_skip(oracle):
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return r


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_perspective_matches_reference(oracle):
    r = _ref_or_skip(oracle)
    out = np.zeros(16, np.float32)
    for fovy, aspect, near, far in [(math.pi / 4, 16 / 9, 1 / 16, synth.FLT_MAX), (1.0, 1.0, 0.1, 100.0), (0.6, 2.39, 1.0, 1000.0)]:
        r.ref_perspective(C.c_float(fovy), C.c_float(aspect), C.c_float(near), C.c_float(far), _vp(out))
        mine = oracle.perspective(fovy, aspect, near, far).reshape(-1)
        assert np.array_equal(mine.view(np.uint32), out.view(np.uint32))
    # the synthetic-scene generator builds the same matrix without touching the oracle
    r.ref_perspective(C.c_float(math.pi / 4), C.c_float(16 / 9), C.c_float(1 / 16), C.c_float(synth.FLT_MAX), _vp(out))
    assert np.array_equal(synth.perspective_inf(math.pi / 4, 16 / 9, 1 / 16).reshape(-1).view(np.uint32), out.view(np.uint32))


def test_inverse_and_mul_match_reference(oracle):
    r = _ref_or_skip(oracle)
    rng = np.random.default_rng(1)
    out = np.zeros(16, np.float32)
    for _ in range(200):
        a = rng.normal(size=16).astype(np.float32)
        b = rng.normal(size=16).astype(np.float32)
        r.ref_mat4_inverse(_vp(a), _vp(out))
        assert np.array_equal(oracle.mat4_inverse(a).reshape(-1).view(np.uint32), out.view(np.uint32))
        r.ref_mat4_mul(_vp(a), _vp(b), _vp(out))
        assert np.array_equal(oracle.mat4_mul(a, b).reshape(-1).view(np.uint32), out.view(np.uint32))


def test_float_to_half_matches_reference_exhaustively_sampled(oracle):
    r = _ref_or_skip(oracle)
    L = oracle.lib()
    rng = np.random.default_rng(2)
    bits = np.concatenate([rng.integers(0, 2 ** 32, size=200000, dtype=np.uint64).astype(np.uint32),
                           np.arange(0x38000000, 0x38000000 + 70000, dtype=np.uint32),   # around the half denormal boundary
                           np.arange(0x477FE000 - 100, 0x477FE000 + 100, dtype=np.uint32)])
    for f in bits.view(np.float32)[:60000]:
        assert L.orc_float_to_half(C.c_float(f)) == r.ref_float_to_half(C.c_float(f))


def test_camera_view_matches_reference(oracle):
    r = _ref_or_skip(oracle)
    eye = np.array([0, 0, 8], np.float32); at = np.zeros(3, np.float32); up = np.array([0, 1, 0], np.float32)
    out = np.zeros(16, np.float32)
    r.ref_camera_view(_vp(eye), _vp(at), _vp(up), _vp(out))
    assert np.array_equal(synth.look_at_view(eye, at).reshape(-1), out)


def test_camera_setup(oracle):
    proj = synth.perspective_inf(math.pi / 4, 16 / 9, 1 / 16)
    view = synth.look_at_view((0, 0, 8), (0, 0, 0))
    cam = oracle.camera_setup(proj, view)
    assert list(cam.camera_position) == [0.0, 0.0, 8.0]
    assert list(cam.camera_front) == [0.0, 0.0, -1.0]
    assert cam.z_near == pytest.approx(1 / 16)
    assert cam.z_far == pytest.approx(6.25e8, rel=1e-5)  # infinite far plane: near / 1e-10


# --------------------------------------------------------------------------------------
# storage formats
# --------------------------------------------------------------------------------------
def test_f16_conversions_against_numpy(oracle):
    L = oracle.lib()
    halves = np.arange(65536, dtype=np.uint16)
    for h in halves[::7]:
        f = np.array([h], np.uint16).view(np.float16).astype(np.float32)[0]
        g = L.orc_f16_to_f32(int(h))
        assert (np.isnan(f) and np.isnan(g)) or np.float32(g).view(np.uint32) == np.float32(f).view(np.uint32)
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.normal(size=20000).astype(np.float32) * 100, rng.uniform(-7e-5, 7e-5, 20000).astype(np.float32),
                           np.array([65504, 65519.99, 65520, 1e9, -1e9, 2 ** -24, 2 ** -25, 2 ** -25 * 1.0001, 0.0, -0.0], np.float32)])
    with np.errstate(over="ignore"):
        expect = vals.astype(np.float16).view(np.uint16)
    for v, e in zip(vals, expect):
        assert L.orc_f32_to_f16(C.c_float(v)) == int(e), v


def test_r11g11b10_properties(oracle):
    L = oracle.lib()
    rgb = np.zeros(3, np.float32)
    prev = -1.0
    for code in range(0, 0x7C0):  # all finite 11-bit codes decode monotonically and round-trip
        L.orc_unpack_r11g11b10(code, _vp(rgb))
        assert rgb[0] > prev
        prev = rgb[0]
        assert L.orc_pack_r11g11b10(C.c_float(rgb[0]), C.c_float(0), C.c_float(0)) == code
        # truncation: anything below the next code packs to this one
        nxt = np.nextafter(np.float32(rgb[0]), np.float32(np.inf))
        assert L.orc_pack_r11g11b10(C.c_float(nxt), C.c_float(0), C.c_float(0)) == code
    assert L.orc_pack_r11g11b10(C.c_float(-1.0), C.c_float(1e30), C.c_float(float("inf"))) == (0x7BF << 11) | (0x3E0 << 22)
    # the numpy packer used by the scene generator agrees with the oracle
    rng = np.random.default_rng(4)
    v = (rng.random((500, 3)) ** 4 * 300).astype(np.float32)
    packed = synth.pack_r11g11b10(v)
    for p, c in zip(packed, v):
        assert int(p) == L.orc_pack_r11g11b10(C.c_float(c[0]), C.c_float(c[1]), C.c_float(c[2]))


def test_srgb_roundtrip(oracle):
    L = oracle.lib()
    last = -1.0
    for v in range(256):
        lin = L.orc_srgb8_to_linear(v)
        assert lin > last
        last = lin
        assert L.orc_linear_to_srgb8(C.c_float(lin)) == v
    assert L.orc_linear_to_srgb8(C.c_float(float("nan"))) == 0
    assert L.orc_linear_to_srgb8(C.c_float(7.0)) == 255


# --------------------------------------------------------------------------------------
# clusterer: the restatement must be CONSERVATIVE (that is the algorithm's defining property)
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,spots", [(16, 0.0), (300, 0.25), (1024, 0.0)])
def test_cluster_is_conservative_and_lighting_agrees_with_brute_force(oracle, n, spots):
    w, h = 320, 180
    scene, cam, lights, prep = common.build_case(oracle, w, h, n, spots)
    clus = oracle.cluster_build(cam, prep)
    hdr, tile, zi, cnt = oracle.deferred_lighting(scene, cam, prep, clus, want_indices=True)
    # brute force: give every pixel every light (all-ones bitmask, full z range)
    full = type(clus)(spots=clus.spots, cull=clus.cull, bitmask=np.full_like(clus.bitmask, 0xFFFFFFFF), range=clus.range.copy())
    if n % 32:
        full.bitmask[..., -1] = np.uint32((1 << (n % 32)) - 1)
    full.range[:, 0] = 0
    full.range[:, 1] = n - 1
    hdr_bf = oracle.deferred_lighting(scene, cam, prep, full)
    # culled lights contribute exactly +0, so the culled result is IDENTICAL to brute force
    assert np.array_equal(hdr, hdr_bf)
    assert cnt.max() <= n
    # indices are in range and sky is marked
    lit = scene.depth != 0
    assert tile[lit].min() >= 0 and tile[lit].max() < 128 * 64
    assert zi[lit].min() >= 0 and zi[lit].max() <= 4095
    assert (tile[~lit] == -1).all()


def test_z_range_against_python_scan(oracle):
    rng = np.random.default_rng(5)
    n = 200
    lo = rng.integers(0, 300, n).astype(np.uint32)
    zr = np.stack([lo, lo + rng.integers(0, 40, n).astype(np.uint32)], -1).astype(np.uint32)
    zr[7] = (0xFFFFFFFF, 0)
    out = np.zeros((512, 2), np.uint32)
    oracle.lib().orc_z_range(_vp(zr), n, 512, _vp(out))
    for z in range(512):
        hit = [i for i in range(n) if zr[i, 0] <= z <= zr[i, 1]]
        exp = (hit[0], hit[-1]) if hit else (0xFFFFFFFF, 0)
        assert tuple(out[z]) == exp


def test_empty_light_list(oracle):
    scene, cam, lights, prep = common.build_case(oracle, 64, 64, 0)
    assert tuple(prep.z_ranges[0]) == (0xFFFFFFFF, 0)
    clus = oracle.cluster_build(cam, prep)
    assert (clus.range[:, 0] == 0xFFFFFFFF).all() and (clus.range[:, 1] == 0).all()
    hdr, _, _, cnt = oracle.deferred_lighting(scene, cam, prep, clus, want_indices=True)
    assert cnt.max() == 0
    sky = scene.depth == 0
    assert np.array_equal(hdr[sky], scene.emissive[sky])


# --------------------------------------------------------------------------------------
# post chain properties
# --------------------------------------------------------------------------------------
def _const16(w, h, rgba):
    return np.broadcast_to(np.array(rgba, np.float16).view(np.uint16), (h, w, 4)).copy()


def test_bloom_filters_preserve_constants(oracle):
    src = _const16(40, 22, (0.5, 2.0, 8.0, -1.25))
    for w, h in [(20, 11), (19, 12)]:
        out = oracle.bloom_downsample(src, (w, h))
        assert np.array_equal(out, _const16(w, h, (0.5, 2.0, 8.0, -1.25)))
    out = oracle.bloom_upsample(src, (80, 44))
    assert np.array_equal(out, _const16(80, 44, (0.5, 2.0, 8.0, -1.25)))
    # feedback: alpha is NOT temporally filtered (mix factor 1), rgb is
    hist = _const16(20, 11, (4.0, 4.0, 4.0, 9.0))
    out = oracle.bloom_downsample(src, (20, 11), hist, 0.25).view(np.float16)
    assert out[..., 3].min() == out[..., 3].max() == np.float16(-1.25)
    assert np.allclose(out[..., 0].astype(np.float32), 4.0 * 0.75 + 0.5 * 0.25, atol=2e-3)


def test_luminance_of_constant_and_clamp(oracle):
    d3 = _const16(60, 34, (0, 0, 0, 1.5))
    lum = oracle.luminance(d3, np.zeros(3, np.float32), 1.0)
    assert lum[0] == pytest.approx(1.5, abs=1e-5) and lum[1] == pytest.approx(2 ** 1.5, rel=1e-5) and lum[2] == pytest.approx(2 ** -1.5, rel=1e-5)
    lum = oracle.luminance(_const16(60, 34, (0, 0, 0, 30.0)), np.zeros(3, np.float32), 1.0)
    assert lum[0] == 2.0  # clamp [-3, 2]
    lum = oracle.luminance(_const16(8, 8, (0, 0, 0, 2.0)), np.array([1.0, 2.0, 0.5], np.float32), 0.5)
    assert lum[0] == pytest.approx(1.5)


def test_frame0_is_black_then_adapts(oracle):
    """Graph buffers start zeroed (render_graph.cpp:2587): avg_inv_lum == 0 on frame 0 => black output."""
    rng = np.random.default_rng(6)
    hdr = common.random_hdr(rng, 128, 72)
    f0 = oracle.hdr_chain(hdr, np.zeros(3, np.float32), None)
    # frame 0 tonemap uses the luminance written THIS frame (lerp from 0): not black, but the
    # threshold used the zero-initialised value
    f1 = oracle.hdr_chain(hdr, f0.lum, f0.d3)
    assert f1.lum[0] != f0.lum[0]
    assert f0.ldr.shape == (72, 128)
    assert ((f0.ldr >> 24) == 0xFF).all()


def test_tonemap_monotone_and_fxaa_flat(oracle):
    ramp = np.linspace(0, 30, 256, dtype=np.float32)
    hdr = synth.pack_r11g11b10(np.stack([ramp, ramp, ramp], -1)[None].repeat(4, 0))
    bloom = np.zeros((1, 64, 4), np.uint16)
    ldr = oracle.tonemap(hdr, bloom, None, 1.0)
    r = (ldr[0] & 0xFF).astype(np.int32)
    assert (np.diff(r) >= 0).all() and r[0] == 0 and r[-1] > 200
    flat = np.full((32, 48), 0xFF336699, np.uint32)
    assert np.array_equal(oracle.fxaa(flat, False), flat)
    out = oracle.fxaa(flat, True)  # decode then the sRGB attachment re-encodes: identity on 8-bit codes
    assert common.rgba8_channel_diff(out, flat).max() <= 1


def test_taa_first_frame_is_identity_within_quantisation(oracle):
    rng = np.random.default_rng(7)
    hdr = common.random_hdr(rng, 64, 40, scale=3.0, hot=0.0)
    col, hist = oracle.taa_resolve(hdr, None, None, None, np.eye(4, dtype=np.float32), 2)
    assert common.max_code_diff_r11g11b10(col, hdr) <= 2
    assert (hist[..., 3] == np.float16(1.0).view(np.uint16)).all()


def test_row_ranges_compose(oracle):
    rng = np.random.default_rng(8)
    hdr = common.random_hdr(rng, 96, 50)
    bloom = common.random_rgba16f(rng, 24, 13, 0, 1)
    full = oracle.tonemap(hdr, bloom, None, 1.0)
    a = oracle.tonemap(hdr, bloom, None, 1.0, rows=(0, 17))
    b = oracle.tonemap(hdr, bloom, None, 1.0, rows=(17, 50))
    assert np.array_equal(full[:17], a[:17]) and np.array_equal(full[17:], b[17:])


# --------------------------------------------------------------------------------------
# light visibility (renderer/scene.cpp:333-358) pinned against the reference's math/frustum.cpp,
# aabb.hpp and simd.hpp (oracle/_ref)
# --------------------------------------------------------------------------------------
def test_frustum_planes_and_box_test_match_reference(oracle):
    r = _ref_or_skip(oracle)
    L = oracle.lib()
    rng = np.random.default_rng(21)
    L.orc_frustum_cull.restype = C.c_int
    r.ref_frustum_cull.restype = C.c_int
    for fovy, aspect, far in [(math.pi / 4, 16 / 9, synth.FLT_MAX), (1.0, 1.0, 200.0), (0.6, 2.39, 1000.0)]:
        proj = oracle.perspective(fovy, aspect, 1 / 16, far)
        for eye in [(0.0, 0.0, 8.0), (3.0, 2.0, -5.0)]:
            cam = oracle.camera_setup(proj, synth.look_at_view(eye, (0.0, 0.0, 0.0)))
            ivp = np.array(list(cam.inv_view_projection), np.float32)
            mine, ref = np.zeros(24, np.float32), np.zeros(24, np.float32)
            L.orc_frustum_planes(_vp(ivp), _vp(mine))
            r.ref_frustum_planes(_vp(ivp), _vp(ref))
            assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
            seen = [0, 0]
            for _ in range(400):
                # boxes around the frustum boundary: a random affine transform of a random static box
                rot = np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(np.float32) * np.float32(rng.uniform(0.5, 2.0))
                rows = np.concatenate([rot, rng.uniform(-60, 60, size=(3, 1)).astype(np.float32)], 1).astype(np.float32).copy()
                lo = -rng.uniform(0.1, 12.0, 3).astype(np.float32)
                hi = rng.uniform(0.0, 12.0, 3).astype(np.float32)
                a_lo, a_hi, b_lo, b_hi = (np.zeros(3, np.float32) for _ in range(4))
                L.orc_transform_aabb(_vp(rows), _vp(lo), _vp(hi), _vp(a_lo), _vp(a_hi))
                r.ref_transform_aabb(_vp(rows), _vp(lo), _vp(hi), _vp(b_lo), _vp(b_hi))
                assert np.array_equal(a_lo.view(np.uint32), b_lo.view(np.uint32)) and np.array_equal(a_hi.view(np.uint32), b_hi.view(np.uint32))
                m, f = L.orc_frustum_cull(_vp(a_lo), _vp(a_hi), _vp(mine)), r.ref_frustum_cull(_vp(b_lo), _vp(b_hi), _vp(ref))
                assert m == f
                seen[m] += 1
            assert seen[0] > 20 and seen[1] > 20, seen
