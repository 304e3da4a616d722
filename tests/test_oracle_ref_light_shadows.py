"""Shadowed positional lights: the oracle's lighting pass with POSITIONAL_LIGHTS_SHADOW (oracle_lighting.c) pinned to the
REFERENCE's own clustering.frag compiled with that define (point.h:45-74, spot.h:51-77, pcf.h:98-99 -> SPIR-V -> C++,
oracle/ref_light_shim.cpp KERNEL=7).  The shader's statements -- the shadow clip transform, the cube reference depth
from shadow[index][0], the products around shadow_falloff -- are the reference's; the comparison sampler itself is a
texture unit's work and is the oracle's restatement of the Vulkan rules on both sides (tests/test_shadow_source_cpu.py
checks that restatement on its own)."""
import os

import numpy as np
import pytest

from tests import common
from tests.test_oracle_ref_light_shaders import compare

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def shadow_case(oracle, w, h, n, spots, res):
    scene, cam, lights, prep = common.build_case(oracle, w, h, n, spots)
    clus = oracle.cluster_build(cam, prep)
    transforms = oracle.shadow_transforms(prep)
    maps = common.make_shadow_maps(prep, res)
    return scene, cam, prep, clus, transforms, maps


@pytest.mark.parametrize("w,h,n,spots,res", [pytest.param(160, 96, 300, 0.25, 32, id="160x96-300-25pct-spots-res32"),
                                             pytest.param(192, 108, 200, 1.0, 64, id="192x108-200-spots-res64"),
                                             pytest.param(128, 128, 64, 0.0, 16, id="128x128-64pt-res16")])
def test_oracle_shadowed_lighting_equals_reference_shader(oracle, w, h, n, spots, res):
    oracle.build()
    k = oracle.ref_light_kernels()
    if k is None or 7 not in k:
        pytest.skip("oracle/_ref lighting shaders not built (no /root/reference on this machine)")
    scene, cam, prep, clus, transforms, maps = shadow_case(oracle, w, h, n, spots, res)
    mine = oracle.deferred_lighting_shadowed(scene, cam, prep, clus, transforms, maps, res)
    ref, d_rgb, c_rgb = oracle.ref_deferred_lighting(scene, cam, prep, clus, shadows=(transforms, maps, res))
    unshadowed = oracle.deferred_lighting(scene, cam, prep, clus)
    _, _, c_plain = oracle.ref_deferred_lighting(scene, cam, prep, clus)
    reached = c_plain.sum(-1) > 0
    changed = (mine != unshadowed)[reached].mean()
    darker = (c_rgb.sum(-1) < c_plain.sum(-1))[reached].mean()
    print(f"clustered light reaches {reached.mean():.3f} of the pixels; shadows change the stored code of {changed:.3f} of those, darken {darker:.3f}")
    assert reached.sum() > 50 and darker > 0.2, "the case must actually shadow pixels"
    compare(mine, ref, "shadowed lit image")
    # all maps absent == the unshadowed pass, bit for bit
    none = oracle.deferred_lighting_shadowed(scene, cam, prep, clus, transforms, [None] * prep.n, res)
    assert np.array_equal(none, unshadowed)


def test_oracle_reproduces_reference_shadow_fixture(oracle):
    """Runs everywhere: the reference shader's shadowed image comes from the committed fixture."""
    f = np.load(os.path.join(GOLDEN, "reflight_shadows_160x96_300.npz"))
    scene, cam, prep, clus, transforms, maps = shadow_case(oracle, 160, 96, 300, 0.25, 32)
    assert np.array_equal(scene.depth, f["depth"]) and np.array_equal(transforms, f["transforms"]), "synthetic case changed: regenerate the fixture"
    compare(oracle.deferred_lighting_shadowed(scene, cam, prep, clus, transforms, maps, 32), f["ref_hdr"], "shadowed image vs fixture")


@pytest.mark.parametrize("w,h,n,spots,res", [pytest.param(160, 96, 300, 0.5, 32, id="160x96-300-50pct-spots-res32"),
                                             pytest.param(192, 108, 200, 1.0, 64, id="192x108-200-spots-res64")])
def test_oracle_wide_pcf_equals_reference_shader(oracle, w, h, n, spots, res):
    """SHADOW_MAP_PCF_KERNEL_WIDE (config "PCFKernelWide"): spot lights filter with the 6 x 6 kernel of pcf.h:7-80."""
    oracle.build()
    k = oracle.ref_light_kernels()
    if k is None or 8 not in k:
        pytest.skip("oracle/_ref lighting shaders not built (no /root/reference on this machine)")
    scene, cam, prep, clus, transforms, maps = shadow_case(oracle, w, h, n, spots, res)
    mine = oracle.deferred_lighting_shadowed(scene, cam, prep, clus, transforms, maps, res, pcf_wide=True)
    ref, _, c_rgb = oracle.ref_deferred_lighting(scene, cam, prep, clus, shadows=(transforms, maps, res, True))
    compare(mine, ref, "wide-PCF lit image")
    narrow = oracle.deferred_lighting_shadowed(scene, cam, prep, clus, transforms, maps, res)
    assert (mine != narrow).mean() > 0.0005, "the wide kernel must change the spot lights' penumbrae"
