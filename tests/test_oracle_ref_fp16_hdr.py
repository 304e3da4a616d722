""""renderTargetFp16" (scene_viewer_application.cpp:880-884): HDR-main / emissive as R16G16B16A16_SFLOAT.  The oracle's
lighting, bloom threshold, tonemap and TAA over an fp16 HDR image, pinned to the REFERENCE's own shaders run on the CPU
with the shims' HDR sampler / blend switched to that format (the shader statements do not change with the format)."""
import numpy as np
import pytest

from tests import common


def f16_image_from_r11(oracle, packed):
    """B10G11R11 image -> the same colours as RGBA16F (alpha 1), plus sub-ulp detail so fp16 carries more than 11 bits."""
    h, w = packed.shape
    L = oracle.lib()
    rgb = np.zeros((h, w, 3), np.float32)
    flat = packed.reshape(-1)
    out = np.zeros(3, np.float32)
    for i in range(flat.size):
        L.orc_unpack_r11g11b10(int(flat[i]), out.ctypes.data_as(oracle.C.c_void_p))
        rgb.reshape(-1, 3)[i] = out
    rng = np.random.default_rng(1)
    rgb *= (1.0 + rng.uniform(-0.004, 0.004, rgb.shape)).astype(np.float32)
    img = np.zeros((h, w, 4), np.float16)
    img[..., :3] = rgb
    img[..., 3] = 1.0
    return np.ascontiguousarray(img).view(np.uint16)


def _skip_without_ref(oracle):
    oracle.build()
    if oracle.ref_post_kernels() is None or oracle.ref_light_kernels() is None:
        pytest.skip("oracle/_ref shaders not built (no /root/reference on this machine)")


def test_fp16_lighting_equals_reference_shaders(oracle):
    _skip_without_ref(oracle)
    scene, cam, lights, prep = common.build_case(oracle, 160, 96, 300, 0.25)
    clus = oracle.cluster_build(cam, prep)
    em16 = f16_image_from_r11(oracle, scene.emissive)
    mine = oracle.deferred_lighting_fp16(scene, cam, prep, clus, em16)
    ref, d_rgb, c_rgb = oracle.ref_deferred_lighting(scene, cam, prep, clus, emissive16=em16)
    sky = scene.depth == 0
    assert np.array_equal(mine[sky], em16[sky]) and np.array_equal(mine[..., 3], em16[..., 3])
    a, b = mine[..., :3].view(np.float16).astype(np.float32), ref[..., :3].view(np.float16).astype(np.float32)
    d = np.abs(mine[..., :3].astype(np.int32) - ref[..., :3].astype(np.int32))  # fp16 codes of non-negative values are ordered
    assert d.max() <= 1 and (d == 0).mean() > 0.995, (int(d.max()), float((d == 0).mean()))
    assert np.abs(a - b).max() <= 0.02 * max(1.0, float(b.max()))
    # shadowed lights take the same path: all maps absent = the unshadowed fp16 image, bit for bit
    none = oracle.deferred_lighting_fp16(scene, cam, prep, clus, em16, shadows=(oracle.shadow_transforms(prep), [None] * prep.n, 32))
    assert np.array_equal(none, mine)


def test_fp16_post_passes_equal_reference_shaders(oracle):
    _skip_without_ref(oracle)
    w, h = 192, 108
    scene, cam, lights, prep = common.build_case(oracle, w, h, 200, 0.0)
    clus = oracle.cluster_build(cam, prep)
    hdr16 = oracle.deferred_lighting_fp16(scene, cam, prep, clus, f16_image_from_r11(oracle, scene.emissive))
    lum = np.array([0.3, 1.1, 0.8], np.float32)
    # K7
    t = oracle.bloom_threshold(hdr16, lum, (w // 2, h // 2))
    t_ref = oracle.ref_bloom_threshold(hdr16, lum, (w // 2, h // 2))
    assert np.array_equal(t[..., :3], t_ref[..., :3])
    assert np.abs(t[..., 3].astype(np.int32) - t_ref[..., 3].astype(np.int32)).max() <= 1  # log2 alpha: libm vs GLM, as for B10G11R11
    # K11
    bloom = oracle.bloom_downsample(t, (w // 4, h // 4))
    assert np.array_equal(oracle.tonemap(hdr16, bloom, lum), oracle.ref_tonemap(hdr16, bloom, lum))
    # K13 (quality 2 with history, and the first frame)
    rng = np.random.default_rng(3)
    mv = np.zeros((h, w, 2), np.float16)
    mv[rng.random((h, w)) < 0.1] = rng.uniform(-2 / w, 2 / w, 2)
    mv = np.ascontiguousarray(mv).view(np.uint16)
    reproj = np.eye(4, dtype=np.float32).T.copy()
    c0, h0 = oracle.taa_resolve(hdr16, scene.depth, mv, None, reproj, 2)
    c0r, h0r = oracle.ref_taa_resolve(hdr16, scene.depth, mv, None, reproj, 2)
    assert np.array_equal(c0, c0r) and np.array_equal(h0, h0r)
    c1, h1 = oracle.taa_resolve(hdr16, scene.depth, mv, h0, reproj, 2)
    c1r, h1r = oracle.ref_taa_resolve(hdr16, scene.depth, mv, h0, reproj, 2)
    assert np.array_equal(c1, c1r) and np.array_equal(h1, h1r)
    # the format matters: the same passes over the B10G11R11 rendition give different codes somewhere
    r11 = oracle.deferred_lighting(scene, cam, prep, clus)
    assert (oracle.tonemap(r11, bloom, lum) != oracle.tonemap(hdr16, bloom, lum)).any()
