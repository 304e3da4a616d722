"""Debug aid: where do the tiled / fast post kernels differ from the oracle?  (GPU box only.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from granite_b200 import capi, harness
from oracle import pyoracle as oracle
from tests import common
from tests.test_oracle_ref_post_shaders import taa_inputs

oracle.build(ref=False)
capi.lib(); capi.init()
rng = np.random.default_rng(1)


def ulp_stats(name, got, ref):
    d = common.f16_ulp_diff(got, ref)
    a = np.abs(got.view(np.float16).astype(np.float32) - ref.view(np.float16).astype(np.float32))
    print(f"{name}: identical {float((d == 0).mean()):.6f}, >1ulp {float((d > 1).mean()):.2e} (max {int(d.max())}), >1ulp and abs>2^-18 {float(((d > 1) & (a > 2.0 ** -18)).mean()):.2e}, max abs {float(a.max()):.3e}")
    return d


for (w, h) in [(1920, 1080)]:
    hdr = common.random_hdr(rng, w, h)
    (tw, th), (dw, dh) = oracle.pyramid_sizes(w, h)[:2]
    for dyn in (False, True):
        lum = np.array([0.3, 2.0 ** 0.3, 2.0 ** -0.3], np.float32) if dyn else None
        ref_t = oracle.bloom_threshold(hdr, lum, (tw, th))
        d0, t = harness.new_rgba16f(dw, dh), harness.new_rgba16f(tw, th)
        harness.bloom_threshold_downsample(harness.to_dev(hdr), harness.to_dev(lum) if dyn else None, d0, t)
        got_t = harness.to_host(t, np.uint16)
        d = ulp_stats(f"fused t dyn={dyn}", got_t, ref_t)
        ys, xs, cs = np.nonzero(d > 1)
        if len(ys):
            print("   worst", [(int(y), int(x), int(c), got_t[y, x].view(np.float16).tolist(), ref_t[y, x].view(np.float16).tolist()) for y, x, c in zip(ys[:4], xs[:4], cs[:4])])
        ulp_stats(f"fused d0 dyn={dyn}", harness.to_host(d0, np.uint16), oracle.bloom_downsample(got_t, (dw, dh)))
for (w, h) in [(1280, 720), (3840, 2160)]:
    r2 = np.random.default_rng(w - h)
    base = r2.integers(0, 256, size=(h // 8 + 1, w // 8 + 1, 4), dtype=np.uint8)
    img = np.kron(base, np.ones((8, 8, 1), np.uint8))[:h, :w].copy()
    img = (img.astype(np.int32) + r2.integers(-6, 7, size=img.shape)).clip(0, 255).astype(np.uint8)
    img32 = np.ascontiguousarray(img).view(np.uint32)[..., 0]
    ref = oracle.fxaa(img32, True)
    out = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    harness.fxaa(harness.to_dev(img32), out, True)
    d = common.rgba8_channel_diff(harness.to_host(out, np.uint32), ref)
    flips = (d > 1).reshape(h, w, 4).any(-1)
    print(f"fxaa {w}x{h}: identical {float((d == 0).mean()):.6f}, flips {int(flips.sum())} ({float(flips.mean()):.2e}), max diff {int(d.max())}")
for (w, h) in [(256, 256), (1280, 720)]:
    r3 = np.random.default_rng(w * 3 + h + 2)
    hdr, depth, mv, hist, reproj = taa_inputs(r3, w, h)
    ref_c, ref_h = oracle.taa_resolve(hdr, depth, mv, hist, reproj, 2)
    oc = torch.zeros((h, w), dtype=torch.int32, device="cuda"); oh = harness.new_rgba16f(w, h)
    harness.taa_resolve(harness.to_dev(hdr), harness.to_dev(depth), harness.to_dev(mv.reshape(h, w, 2)).view(torch.int32).reshape(h, w), harness.to_dev(hist), reproj, 2, oc, oh)
    got_c, got_h = harness.to_host(oc, np.uint32), harness.to_host(oh, np.uint16)
    d = ulp_stats(f"taa {w}x{h} history", got_h, ref_h)
    dc = np.max([np.abs(x - y) for x, y in zip(common.r11g11b10_codes(got_c), common.r11g11b10_codes(ref_c))], axis=0)
    print(f"   colour: identical {float((dc == 0).mean()):.6f}, >1 code {float((dc > 1).mean()):.2e}, max {int(dc.max())}")
    a = np.abs(got_h.view(np.float16).astype(np.float32) - ref_h.view(np.float16).astype(np.float32))
    ys, xs, cs = np.nonzero((d > 1) & (a > 2.0 ** -18))
    mvf = mv.view(np.float16).reshape(h, w, 2)
    print("   worst", [(int(y), int(x), int(c), got_h[y, x].view(np.float16).tolist(), ref_h[y, x].view(np.float16).tolist(), mvf[y, x].tolist()) for y, x, c in zip(ys[:5], xs[:5], cs[:5])])
    if len(ys):
        print("   bad rows hist", np.bincount(ys // max(h // 8, 1)).tolist(), "cols hist", np.bincount(xs // max(w // 8, 1)).tolist(), "chan hist", np.bincount(cs, minlength=4).tolist(),
              "frac with mv!=0", float((mvf[ys, xs] != 0).any(-1).mean()))
