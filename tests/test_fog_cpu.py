"""Volumetric fog, accumulation pass (volumetric_fog.cpp:236-254, fog_accumulate.comp): the oracle pinned
to the reference's own shader run on the CPU, the kernel of granite_b200/csrc/grb_fog.cu compiled for the CPU bit for bit
against the oracle, and the physics of the recurrence."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_density(w, h, d, seed=0):
    """(d, h, w, 4) RGBA16F froxel grid: smooth in-scattered light with a few bright froxels, optical depth 0 .. 0.4 with
    empty regions and one dense wall."""
    rng = np.random.default_rng(seed + w * 3 + d)
    zz, yy, xx = np.mgrid[0:d, 0:h, 0:w]
    light = np.stack([0.4 + 0.3 * np.sin(xx * 0.3 + zz * 0.2), 0.3 + 0.2 * np.cos(yy * 0.25), 0.2 + 0.1 * np.sin(zz * 0.5)], -1)
    light = light * rng.uniform(0.5, 1.5, (d, h, w, 1))
    light[rng.random((d, h, w)) < 0.002] = (30.0, 20.0, 10.0)
    a = rng.uniform(0.0, 0.4, (d, h, w))
    a[rng.random((d, h, w)) < 0.3] = 0.0
    a[d // 2, :, : w // 3] = 6.0
    vol = np.concatenate([light, a[..., None]], -1).astype(np.float16)
    return np.ascontiguousarray(vol).view(np.uint16)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libemu_fog.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{cuda}/include",
           os.path.join(ROOT, "tests", "cpp", "emulate_fog.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(out)


def _f16(a):
    return a.view(np.float16).astype(np.float32)


@pytest.mark.parametrize("w,h,d", [(40, 23, 16), (33, 17, 7), (160, 92, 64)])
def test_oracle_fog_accumulate_equals_reference_shader(oracle, w, h, d):
    oracle.build()
    k = oracle.ref_post_kernels()
    if k is None or 27 not in k:
        pytest.skip("oracle/_ref shaders not built (no /root/reference on this machine)")
    light = make_density(w, h, d)
    mine, ref = oracle.fog_accumulate(light), oracle.ref_fog_accumulate(light)
    # the recurrence's only transcendental is exp2: glibc here, GLM's exp(x ln 2) in the generated code
    diff = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999, (int(diff.max()), float((diff == 0).mean()))


@pytest.mark.parametrize("w,h,d", [(40, 23, 16), (33, 17, 7), (65, 9, 3), (160, 92, 64)])
def test_kernel_source_equals_oracle(emu, oracle, w, h, d):
    light = make_density(w, h, d, seed=5)
    got = np.zeros_like(light)
    emu.emu_fog_accumulate(light.ctypes.data_as(C.c_void_p), w, h, d, got.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, oracle.fog_accumulate(light))


def test_fog_accumulate_physics(oracle):
    """Transmittance falls monotonically and equals exp2(-sum of blurred depths); light only grows; an empty grid gives
    (0, 0, 0, 1); behind an opaque wall nothing more is added."""
    w, h, d = 24, 12, 20
    empty = np.zeros((d, h, w, 4), np.float16).view(np.uint16)
    out = _f16(oracle.fog_accumulate(empty))
    assert (out[..., :3] == 0).all() and (out[..., 3] == 1).all()
    light = make_density(w, h, d, seed=9)
    fog = _f16(oracle.fog_accumulate(light))
    assert (np.diff(fog[..., 3], axis=0) <= 1e-6).all(), "transmittance never rises along the ray"
    assert (np.diff(fog[..., :3], axis=0) >= -2e-3 * np.maximum(fog[1:, ..., :3], 1)).all(), "accumulated light never falls (fp16 rounding aside)"
    wall = np.zeros((d, h, w, 4), np.float16)
    wall[..., :3] = 1.0
    wall[..., 3] = 0.01
    wall[5] = (1.0, 1.0, 1.0, 60.0)
    fog = _f16(oracle.fog_accumulate(wall.view(np.uint16)))
    assert (fog[8:, ..., 3] < 1e-6).all() and np.allclose(fog[8], fog[-1], atol=2e-3 * fog[8].max())
    # uniform medium: the 17 weights sum to 1 (1.375 / 1.375), so the depth after k slices is k * a
    uni = np.zeros((d, h, w, 4), np.float16)
    uni[..., 3] = 0.125
    t = _f16(oracle.fog_accumulate(uni.view(np.uint16)))[:, h // 2, w // 2, 3]
    assert np.allclose(t, np.exp2(-0.125 * np.arange(1, d + 1)), rtol=2e-3)


# ---- light-density pass (fog_light_density.comp, base variant) ----
def fog_case(oracle, w=40, h=23, d=16, n=300, spots=0.25):
    from tests import common

    cam, lights, prep = common.build_lights_case(oracle, 16.0 / 9.0, n, spots)
    clus = oracle.cluster_build(cam, prep)
    fp = oracle.fog_params(w, h, d, z_range=80.0, density=0.5, in_scatter=1.25, dither_offset=1)
    lut = np.random.default_rng(17).integers(0, 2 ** 32, (3, 128, 128), dtype=np.uint64).astype(np.uint32)
    return cam, prep, clus, fp, lut


DIR_COLOR, DIR_DIRECTION = (6.0, 5.5, 4.5), tuple(np.array([0.3, 0.8, 0.5]) / np.linalg.norm([0.3, 0.8, 0.5]))


def test_fog_density_kernel_source_equals_oracle(emu, oracle):
    from granite_b200 import capi

    cam, prep, clus, fp, lut = fog_case(oracle)
    ref = oracle.fog_light_density(fp, cam, prep, clus, DIR_COLOR, DIR_DIRECTION, lut)
    g = capi.GrbFogParameters(fp.width, fp.height, fp.depth, fp.dither_offset, fp.slice_z_log2_scale, fp.density_mod, fp.in_scatter_strength)
    gcam = capi.GrbCamera()
    for name in ("view", "view_projection", "inv_view_projection", "camera_position", "camera_front"):
        getattr(gcam, name)[:] = list(getattr(cam, name))
    gp = capi.GrbClusterParameters()
    for name, _t in capi.GrbClusterParameters._fields_:
        v = getattr(prep.params, name)
        if hasattr(v, "__len__"):
            getattr(gp, name)[:] = list(v)
        else:
            setattr(gp, name, v)
    keep = [np.ascontiguousarray(prep.records), np.ascontiguousarray(prep.type_mask, np.uint32), np.ascontiguousarray(clus.bitmask, np.uint32),
            np.ascontiguousarray(clus.range, np.uint32)]
    buf = capi.GrbClusterBuffers()
    buf.lights, buf.type_mask, buf.bitmask, buf.cluster_range = [k.ctypes.data for k in keep]
    ext = oracle.fog_slice_extents(fp)
    proj, inv_proj = np.array(list(cam.projection), np.float32), np.array(list(cam.inv_projection), np.float32)
    dc, dd = np.array(DIR_COLOR, np.float32), np.array(DIR_DIRECTION, np.float32)
    got = np.zeros_like(ref)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    emu.emu_fog_light_density(C.byref(g), C.byref(gcam), p(proj), p(inv_proj), C.byref(gp), C.byref(buf), p(dc), p(dd), p(ext), p(lut), p(got))
    assert np.array_equal(got, ref)
    from tests import common

    cam0, _, prep0 = common.build_lights_case(oracle, 16.0 / 9.0, 0)
    dark = oracle.fog_light_density(fp, cam0, prep0, oracle.cluster_build(cam0, prep0), DIR_COLOR, DIR_DIRECTION, lut)
    reached = (ref[..., :3] != dark[..., :3]).any(-1).mean()  # froxels a positional light adds to
    print(f"positional lights reach {reached:.4f} of the froxels")
    assert 0.003 < reached < 0.98 and np.isfinite(_f16(ref)).all() and np.array_equal(ref[..., 3], dark[..., 3])


def test_fog_density_properties(oracle):
    """No positional lights: in-scatter = strength * colour * (0.55 - 0.45 VoL) within the phase function's range; the albedo
    grows with the slice thickness along z and with the ray's obliquity towards the screen corners."""
    from tests import common

    cam, lights, prep = common.build_lights_case(oracle, 16.0 / 9.0, 0)
    clus = oracle.cluster_build(cam, prep)
    fp = oracle.fog_params(32, 18, 12, in_scatter=2.0)
    lut = np.full((1, 128, 128), 0x00007F7F, np.uint32)  # dither (0.498 - 0.5, 0.498 - 0.5, 0): practically none
    out = _f16(oracle.fog_light_density(fp, cam, prep, clus, DIR_COLOR, DIR_DIRECTION, lut))
    for c in range(3):
        lo, hi = 2.0 * DIR_COLOR[c] * 0.1, 2.0 * DIR_COLOR[c] * 1.0
        assert (out[..., c] >= lo * 0.999).all() and (out[..., c] <= hi * 1.001).all()
    a = out[..., 3]
    assert (np.diff(a[::4, 9, 16]) > 0).all(), "slices get thicker with distance (the shader reads one extent per group of four slices)"
    assert a[5, 0, 0] > a[5, 9, 16] and a[5, 17, 31] > a[5, 9, 16], "oblique rays cross more medium"
    ext = oracle.fog_slice_extents(fp)
    assert abs(ext.sum() - 80.0) < 1e-2, "the slices tile [0, z_range]"


@pytest.mark.parametrize("n,spots", [(300, 0.25), (64, 1.0), (0, 0.0)])
def test_oracle_fog_density_equals_reference_shader(oracle, n, spots):
    """fog_light_density.comp compiled as the forward renderer compiles it for an unshadowed scene (STAGE_COMPUTE,
    RENDERER_FORWARD, POSITIONAL_LIGHTS, CLUSTERER_BINDLESS; no FOG_REGIONS / TEMPORAL_REPROJECTION / FLOOR_LIGHTING) and run on
    the CPU.  exp2 / sqrt are libm's here and GLM's there: stored fp16 values within one code."""
    oracle.build()
    k = oracle.ref_light_kernels()
    if k is None or 9 not in k:
        pytest.skip("oracle/_ref shaders not built (no /root/reference on this machine)")
    cam, prep, clus, fp, lut = fog_case(oracle, n=n, spots=spots)
    mine = oracle.fog_light_density(fp, cam, prep, clus, DIR_COLOR, DIR_DIRECTION, lut)
    ref = oracle.ref_fog_light_density(fp, cam, prep, clus, DIR_COLOR, DIR_DIRECTION, lut)
    diff = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.995, (int(diff.max()), float((diff == 0).mean()))
