"""The oracle's deferred lighting (oracle_lighting.c: K6 directional + K5 clustered point / spot lights)
pinned to the REFERENCE's own fragment shaders.

assets/shaders/lights/{directional,clustering}.frag (+ clusterer_bindless.h, point.h, spot.h, pbr.h,
lighting.h) -> SPIR-V (the reference's vendored glslang) -> C++ (its vendored spirv-cross, `--cpp`) ->
executed per pixel on the CPU (`make -C oracle ref-shaders`, oracle/ref_light_shim.cpp: what the shim
supplies -- G-buffer texel decode, the interpolated vClip, a subgroup of one invocation -- is listed in
its header).  The fp32 colours of the two draws are then blended into B10G11R11 by the oracle's store
rule and compared with oracle_deferred_lighting's image:

  * cluster addressing (tile, Z slice, mask range) is integer: any disagreement changes which lights
    a pixel sums and shows up as a gross difference -- none is tolerated;
  * the BRDF is fp32 with sqrt / pow / division whose association the shader compiler is free to
    choose, so the stored 11/11/10-bit codes may differ by 1 on a small fraction of the pixels.
"""
import os

import numpy as np
import pytest

from tests import common

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _light_or_skip(oracle):
    oracle.build()
    if oracle.ref_light_kernels() is None:
        pytest.skip("oracle/_ref lighting shaders not built (no /root/reference on this machine)")


def compare(mine, ref, what):
    codes_a, codes_b = common.r11g11b10_codes(mine), common.r11g11b10_codes(ref)
    d = np.max([np.abs(a - b) for a, b in zip(codes_a, codes_b)], axis=0)
    assert d.max() <= 1, f"{what}: {int((d > 1).sum())} pixels differ by more than one code (max {int(d.max())})"
    assert (d == 0).mean() > 0.999, f"{what}: only {(d == 0).mean():.5f} of the pixels identical"


@pytest.mark.parametrize("w,h,n,spots", [pytest.param(256, 256, 16, 0.0, id="C1-256-16pt"), pytest.param(160, 96, 300, 0.25, id="160x96-300-25pct-spots"),
                                         pytest.param(192, 108, 1024, 0.0, id="192x108-1024pt")])
def test_oracle_lighting_equals_reference_shaders(oracle, w, h, n, spots):
    _light_or_skip(oracle)
    scene, cam, lights, prep = common.build_case(oracle, w, h, n, spots)
    clus = oracle.cluster_build(cam, prep)
    mine, tile, zidx, cnt = oracle.deferred_lighting(scene, cam, prep, clus, want_indices=True)
    ref, d_rgb, c_rgb = oracle.ref_deferred_lighting(scene, cam, prep, clus)
    assert cnt.max() >= 2 and (c_rgb.sum(-1) > 0).mean() > 0.005 and (d_rgb.sum(-1) > 0).mean() > 0.5, "the case must actually light pixels"
    print(f"clustered light reaches {(c_rgb.sum(-1) > 0).mean():.3f} of the pixels, up to {int(cnt.max())} lights per pixel")
    sky = scene.depth == 0
    assert np.array_equal(mine[sky], ref[sky]) and np.array_equal(ref[sky], scene.emissive[sky])
    compare(mine, ref, "lit image")


def test_oracle_reproduces_reference_lighting_fixture(oracle):
    """Runs everywhere: the reference shaders' image comes from the committed fixture."""
    f = np.load(os.path.join(GOLDEN, "reflight_160x96_300_25pct_spots.npz"))
    scene, cam, lights, prep = common.build_case(oracle, 160, 96, 300, 0.25)
    clus = oracle.cluster_build(cam, prep)
    assert np.array_equal(scene.depth, f["depth"]) and np.array_equal(scene.albedo, f["albedo"]), "synthetic scene generator changed: regenerate the fixture"
    compare(oracle.deferred_lighting(scene, cam, prep, clus), f["ref_hdr"], "lit image vs fixture")
