"""CUDA kernels against vectors produced by the REFERENCE's own shaders (tests/golden/ref*.npz, made in
the build container by running the reference's GLSL through its vendored glslang + spirv-cross on the
CPU: tests/golden/make_ref_*_golden.py).  These run on the GPU box, where /root/reference does not
exist: every stage is fed the fixture's input and compared with the fixture's output."""
import os

import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cuda_post_chain_vs_reference_shader_fixture(cuda):
    import torch

    from granite_b200 import harness

    f = np.load(os.path.join(GOLDEN, "refpost_chain_270x135.npz"))
    hdr = f["hdr"]
    h, w = hdr.shape
    dev = harness.to_dev
    for frame in range(2):
        g = lambda k: f[f"f{frame}_{k}"]
        lum_in = f["lum_in"] if frame == 0 else f["f0_lum"]
        hist = None if frame == 0 else f["f0_d3"]
        # K7 threshold (bloom_threshold.comp): rgb exact, alpha = log2 within 1 fp16 ulp
        t = harness.new_rgba16f(*g("t").shape[:2][::-1])
        harness.bloom_threshold(dev(hdr), dev(lum_in), t)
        got = harness.to_host(t, np.uint16)
        assert np.array_equal(got[..., :3], g("t")[..., :3]) and common.f16_ulp_diff(got[..., 3], g("t")[..., 3]).max() <= 1
        # K8 / K9 (bloom_downsample.comp, bloom_upsample.comp): no transcendental, bit for bit
        for src, dst, hst in (("t", "d0", None), ("d0", "d1", None), ("d1", "d2", None), ("d2", "d3", hist)):
            out = harness.new_rgba16f(*g(dst).shape[:2][::-1])
            harness.bloom_downsample(dev(g(src)), out, dev(hst) if hst is not None else None, float(np.float32(1.0 - 0.001 ** (1 / 60))))
            assert np.array_equal(harness.to_host(out, np.uint16), g(dst)), f"frame {frame}: {dst}"  # 270x135: no exact 2:1 step, generic kernels
        for src, dst in (("d3", "u2"), ("u2", "u1"), ("u1", "u0")):
            out = harness.new_rgba16f(*g(dst).shape[:2][::-1])
            harness.bloom_upsample(dev(g(src)), out)
            assert np.array_equal(harness.to_host(out, np.uint16), g(dst)), f"frame {frame}: {dst}"
        # K10 luminance.comp: the log-average exact, its exp2 outputs within a few ulps
        lum_t = dev(lum_in.copy())
        harness.luminance(dev(g("d3")), lum_t, float(np.float32(1.0 - 0.5 ** (1 / 60))))
        lum = lum_t.cpu().numpy()
        assert lum.view(np.uint32)[0] == g("lum").view(np.uint32)[0]
        assert common.f32_ulp_diff(lum[1:], g("lum")[1:]).max() <= 4
        # K11 tonemap.frag: 1 LSB
        ldr = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        harness.tonemap(dev(hdr), dev(g("u0")), dev(g("lum")), ldr, exposure=1.0)
        d = common.rgba8_channel_diff(harness.to_host(ldr, np.uint32), g("ldr"))
        assert d.max() <= 1 and (d == 0).mean() > 0.999


def test_cuda_aa_vs_reference_shader_fixture(cuda):
    import torch

    from granite_b200 import harness

    g = np.load(os.path.join(GOLDEN, "refpost_aa_128x80.npz"))
    h, w = g["ldr"].shape
    for srgb, key in ((True, "fxaa_srgb"), (False, "fxaa_unorm")):
        out = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        harness.fxaa(harness.to_dev(g["ldr"]), out, srgb)
        d = common.rgba8_channel_diff(harness.to_host(out, np.uint32), g[key])
        assert ((d > 1).reshape(h, w, 4).any(-1)).mean() <= 2e-4 and (d == 0).mean() > 0.995, key  # see test_fxaa about branch flips
    hdr_t, depth_t = harness.to_dev(g["hdr"]), harness.to_dev(g["depth"])
    mv_t = harness.to_dev(g["mv"].reshape(h, w, 2)).view(torch.int32).reshape(h, w)
    for q in (0, 1, 2):
        oc = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        oh = harness.new_rgba16f(w, h)
        harness.taa_resolve(hdr_t, depth_t, mv_t, harness.to_dev(g["hist"]), g["reproj"], q, oc, oh)
        got_c, got_h = harness.to_host(oc, np.uint32), harness.to_host(oh, np.uint16)
        assert np.array_equal(got_c, g[f"taa_q{q}_color"]) and np.array_equal(got_h, g[f"taa_q{q}_history"])
    oc = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    oh = harness.new_rgba16f(w, h)
    harness.taa_resolve(hdr_t, None, None, None, None, 2, oc, oh)
    assert np.array_equal(harness.to_host(oc, np.uint32), g["taa_first_color"]) and np.array_equal(harness.to_host(oh, np.uint16), g["taa_first_history"])


def test_cuda_pq10_vs_reference_shader_fixture(cuda):
    """pq10_encode.frag: the kernel evaluates the two pow() per channel with lg2 / ex2, so a code may round the other
    way at an exact .5 of the 10-bit scale: 1 code, rarely."""
    import torch

    from granite_b200 import harness

    p = np.load(os.path.join(GOLDEN, "refpost_pq10_96x64.npz"))
    h, w = p["hdr"].shape
    out = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    harness.pq10_encode(harness.to_dev(p["hdr"]), harness.to_dev(p["ui"]), p["primary_conversion"], 500.0, 400.0, 1000.0, out)
    d = common.a2b10g10r10_channel_diff(harness.to_host(out, np.uint32), p["pq10"])
    assert d.max() <= 1 and (d == 0).mean() > 0.99


def test_cuda_lighting_vs_reference_shader_fixture(cuda, oracle):
    """HDR-main lit by the CUDA kernel vs the image the reference's directional.frag + clustering.frag
    produce for the same seeded scene (the cluster structure is built on the GPU too)."""
    from granite_b200 import harness

    from tests.test_gpu_parity import _cluster

    f = np.load(os.path.join(GOLDEN, "reflight_160x96_300_25pct_spots.npz"))
    scene, cam, lights, prep = common.build_case(oracle, 160, 96, 300, 0.25)
    assert np.array_equal(scene.depth, f["depth"]) and np.array_equal(scene.albedo, f["albedo"]), "scene generator changed: regenerate the fixture"
    dev, gcam = _cluster(cuda, oracle, cam, prep)
    gb = harness.GBufferDevice(scene)
    hdr = gb.emissive.clone()
    harness.deferred_lighting(gb, gcam, dev, hdr)
    got = harness.to_host(hdr, np.uint32)
    d = np.max([np.abs(a - b) for a, b in zip(common.r11g11b10_codes(got), common.r11g11b10_codes(f["ref_hdr"]))], axis=0)
    print(f"lighting vs reference shaders: identical {float((d == 0).mean()):.5f}, max code difference {int(d.max())}")
    assert d.max() <= 1 and (d == 0).mean() > 0.97
