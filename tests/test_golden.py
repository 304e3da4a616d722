"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py).
CPU: the oracle still reproduces them bit for bit (guards against oracle drift).
GPU: the CUDA path reproduces them from the stored inputs WITHOUT the oracle in the loop."""
import ctypes as C
import os

import numpy as np
import pytest

from granite_b200 import synth
from tests import common

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_CASES = [("frame_96x64_40lights", 40, 0.25, 2), ("frame_c1_256x256_16lights", 16, 0.0, 1)]


def _load(name):
    return dict(np.load(os.path.join(HERE, name + ".npz")))


def _camera(oracle_mod, g):
    cam = oracle_mod.Camera()
    for k in ("projection", "view", "view_projection", "inv_projection", "inv_view", "inv_view_projection", "camera_position", "camera_front"):
        getattr(cam, k)[:] = g["cam_" + k].tolist()
    cam.z_near, cam.z_far = float(g["cam_z"][0]), float(g["cam_z"][1])
    return cam


@pytest.mark.parametrize("name,n,spots,frames", FRAME_CASES)
def test_oracle_reproduces_frame_fixture(oracle, name, n, spots, frames):
    g = _load(name)
    h, w = g["depth"].shape
    scene = synth.make_scene(w, h)
    for k in ("albedo", "normal", "pbr", "depth", "emissive"):
        assert np.array_equal(getattr(scene, k), g[k]), f"synthetic generator changed: {k}"
    cam = oracle.camera_setup(scene.projection, scene.view)
    assert np.array_equal(np.array(list(cam.inv_view_projection), np.float32), g["cam_inv_view_projection"])
    prep = oracle.prepare_lights(cam, synth.make_lights(n, spot_fraction=spots, aspect=w / h))
    assert prep.records[:max(n, 1)].tobytes() == g["records"].tobytes()
    assert np.array_equal(prep.z_ranges, g["z_ranges"])
    clus = oracle.cluster_build(cam, prep)
    assert np.array_equal(clus.bitmask, g["bitmask"]) and np.array_equal(clus.range, g["cluster_range"])
    hdr, tile, zi, cnt = oracle.deferred_lighting(scene, cam, prep, clus, want_indices=True)
    assert np.array_equal(hdr, g["hdr"]) and np.array_equal(tile, g["tile_index"]) and np.array_equal(zi, g["z_index"])
    lum, d3 = np.zeros(3, np.float32), None
    for i in range(frames):
        f = oracle.hdr_chain(hdr, lum, d3)
        lum, d3 = f.lum, f.d3
        for k in ("t", "d0", "d3", "u0", "ldr"):
            assert np.array_equal(getattr(f, k), g[f"f{i}_{k}"]), (i, k)
        assert np.array_equal(f.lum.view(np.uint32), g[f"f{i}_lum"].view(np.uint32))
    assert np.array_equal(oracle.fxaa(f.ldr, True), g["fxaa_srgb"])


def test_oracle_reproduces_taa_fixture(oracle):
    g = _load("taa_80x48")
    for q in (0, 1, 2):
        c0, h0 = oracle.taa_resolve(g["hdr"], g["depth"], g["mv"], None, g["reproj"], q)
        c1, h1 = oracle.taa_resolve(g["hdr"], g["depth"], g["mv"], h0, g["reproj"], q)
        assert np.array_equal(c0, g[f"q{q}_color0"]) and np.array_equal(h0, g[f"q{q}_hist0"])
        assert np.array_equal(c1, g[f"q{q}_color1"]) and np.array_equal(h1, g[f"q{q}_hist1"])


# ----------------------------------------------------------------------------------- GPU replay
class _Cam:
    pass


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,spots,frames", FRAME_CASES)
def test_cuda_reproduces_frame_fixture(cuda, name, n, spots, frames):
    import torch

    from granite_b200 import capi, harness

    g = _load(name)
    h, w = g["depth"].shape
    cam = _Cam()
    for k in ("view", "view_projection", "inv_view_projection", "camera_position", "camera_front"):
        setattr(cam, k, g["cam_" + k].tolist())
    cam.z_near, cam.z_far = float(g["cam_z"][0]), float(g["cam_z"][1])
    gcam = harness.camera_struct(cam)
    params = capi.GrbClusterParameters.from_buffer_copy(g["params"].tobytes())
    dev = harness.ClusterDevice(g["records"], g["model"], g["type_mask"], g["z_ranges"], params, synth.CLUSTER_RES)
    dev.build(gcam)
    got = dev.download()
    assert np.array_equal(got.bitmask, g["bitmask"]) and np.array_equal(got.range, g["cluster_range"])

    scene = synth.Scene(w, h, None, None, g["albedo"], g["normal"], g["pbr"], g["depth"], g["emissive"],
                        dir_color=tuple(g["dir_color"].tolist()), dir_direction=tuple(g["dir_direction"].tolist()))
    gb = harness.GBufferDevice(scene)
    hdr = gb.emissive.clone()
    harness.deferred_lighting(gb, gcam, dev, hdr)
    assert common.max_code_diff_r11g11b10(harness.to_host(hdr, np.uint32), g["hdr"]) <= 1
    out_t = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    out_z = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    dimg = capi.image(gb.depth, capi.FORMAT_D32_SFLOAT)
    capi.check(capi.lib().grb_debug_cluster_indices(C.byref(dimg), C.byref(gcam), C.byref(dev.params), C.c_void_p(out_t.data_ptr()),
                                                    C.c_void_p(out_z.data_ptr()), capi.rows(), capi.stream_ptr()))
    assert np.array_equal(out_t.cpu().numpy(), g["tile_index"]) and np.array_equal(out_z.cpu().numpy(), g["z_index"])

    # post chain replayed from the FIXTURE's hdr so every level is comparable bit for bit
    sz = [g[f"f0_{k}"].shape[:2][::-1] for k in ("t", "d0", "d1", "d2", "d3")]
    hdr_t = harness.to_dev(g["hdr"])
    lum_t = torch.zeros(3, dtype=torch.float32, device="cuda")
    lv = {k: harness.new_rgba16f(*s) for k, s in zip(("t", "d0", "d1", "d2", "d3"), sz)}
    up = {"u2": harness.new_rgba16f(*sz[3]), "u1": harness.new_rgba16f(*sz[2]), "u0": harness.new_rgba16f(*sz[1])}
    d3_prev = None
    ldr = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    for i in range(frames):
        harness.bloom_threshold(hdr_t, lum_t, lv["t"])
        harness.bloom_downsample(lv["t"], lv["d0"])
        harness.bloom_downsample(lv["d0"], lv["d1"])
        harness.bloom_downsample(lv["d1"], lv["d2"])
        d3_new = harness.new_rgba16f(*sz[4])
        harness.bloom_downsample(lv["d2"], d3_new, d3_prev, float(np.float32(1.0 - 0.001 ** (1 / 60))))
        harness.luminance(d3_new, lum_t, float(np.float32(1.0 - 0.5 ** (1 / 60))))
        harness.bloom_upsample(d3_new, up["u2"])
        harness.bloom_upsample(up["u2"], up["u1"])
        harness.bloom_upsample(up["u1"], up["u0"])
        harness.tonemap(hdr_t, up["u0"], lum_t, ldr)
        t = harness.to_host(lv["t"], np.uint16)
        assert common.f16_ulp_diff(t, g[f"f{i}_t"]).max() <= 1
        assert common.f16_ulp_diff(harness.to_host(d3_new, np.uint16), g[f"f{i}_d3"]).max() <= 2
        assert abs(float(lum_t.cpu()[0]) - float(g[f"f{i}_lum"][0])) < 1e-4
        assert common.rgba8_channel_diff(harness.to_host(ldr, np.uint32), g[f"f{i}_ldr"]).max() <= 1
        d3_prev = d3_new


@pytest.mark.gpu
def test_cuda_reproduces_taa_fixture(cuda):
    import torch

    from granite_b200 import harness

    g = _load("taa_80x48")
    h, w = g["depth"].shape
    hdr_t = harness.to_dev(g["hdr"])
    depth_t = harness.to_dev(g["depth"])
    mv_t = harness.to_dev(g["mv"].reshape(h, w, 2)).view(torch.int32).reshape(h, w)
    for q in (0, 1, 2):
        oc = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        oh = harness.new_rgba16f(w, h)
        harness.taa_resolve(hdr_t, None, None, None, None, q, oc, oh)
        assert np.array_equal(harness.to_host(oc, np.uint32), g[f"q{q}_color0"])
        assert np.array_equal(harness.to_host(oh, np.uint16), g[f"q{q}_hist0"])
        oc1 = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        oh1 = harness.new_rgba16f(w, h)
        harness.taa_resolve(hdr_t, depth_t, mv_t, oh, g["reproj"], q, oc1, oh1)
        assert np.array_equal(harness.to_host(oc1, np.uint32), g[f"q{q}_color1"])
        assert np.array_equal(harness.to_host(oh1, np.uint16), g[f"q{q}_hist1"])
