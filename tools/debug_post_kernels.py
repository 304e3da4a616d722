"""Debug aid: where do the tiled / fast post kernels differ from the oracle?  (GPU box only.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from granite_b200 import capi, harness
from oracle import pyoracle as oracle
from tests import common

oracle.build(ref=False)
capi.lib(); capi.init()


def report(name, got, ref, chan_last=True):
    bad = (got != ref)
    if bad.ndim == 3:
        bad = bad.any(-1)
    ys, xs = np.nonzero(bad)
    print(f"{name}: {bad.sum()} / {bad.size} texels differ", end="")
    if len(ys):
        print(f"; rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; first {list(zip(ys[:6].tolist(), xs[:6].tolist()))}; "
              f"x%32 hist top {np.bincount(xs % 32).argsort()[-4:][::-1].tolist()} y%16 top {np.bincount(ys % 16).argsort()[-4:][::-1].tolist()}")
        y, x = ys[0], xs[0]
        print("   got", got[y, x], "ref", ref[y, x])
    else:
        print()


rng = np.random.default_rng(1)
for (w_in, h_in, w, h) in [(128, 128, 64, 64), (192, 64, 96, 32), (960, 540, 480, 270)]:
    src = common.random_rgba16f(rng, w_in, h_in)
    ref = oracle.bloom_downsample(src, (w, h))
    out = harness.new_rgba16f(w, h)
    harness.bloom_downsample(harness.to_dev(src), out)
    report(f"down {w_in}x{h_in}", harness.to_host(out, np.uint16), ref)
for (w_in, h_in, w, h) in [(30, 17, 60, 34), (96, 32, 192, 64), (480, 270, 960, 540)]:
    src = common.random_rgba16f(rng, w_in, h_in)
    ref = oracle.bloom_upsample(src, (w, h))
    out = harness.new_rgba16f(w, h)
    harness.bloom_upsample(harness.to_dev(src), out)
    report(f"up {w_in}x{h_in}", harness.to_host(out, np.uint16), ref)
for (w, h) in [(1920, 1080), (3840, 2160)]:
    hdr = common.random_hdr(rng, w, h)
    (tw, th), (dw, dh) = oracle.pyramid_sizes(w, h)[:2]
    t = harness.new_rgba16f(tw, th)
    harness.bloom_threshold(harness.to_dev(hdr), None, t)
    ref = oracle.bloom_downsample(harness.to_host(t, np.uint16), (dw, dh))
    d0 = harness.new_rgba16f(dw, dh)
    harness.bloom_threshold_downsample(harness.to_dev(hdr), None, d0)
    report(f"fused {w}x{h}", harness.to_host(d0, np.uint16), ref)
w, h = 256, 256
img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
img = np.ascontiguousarray(img).view(np.uint32).reshape(h, w)
for srgb in (False, True):
    ref = oracle.fxaa(img, srgb)
    out = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    harness.fxaa(harness.to_dev(img), out, srgb)
    got = harness.to_host(out, np.uint32)
    d = common.rgba8_channel_diff(got, ref)
    print("fxaa srgb", srgb, "max diff", d.max(), "frac nonzero", (d > 0).mean(), "frac >1", (d > 1).mean())
    ys, xs, cs = np.nonzero(d.reshape(h, w, 4) > 1)
    print("   first bad", list(zip(ys[:5].tolist(), xs[:5].tolist(), cs[:5].tolist())), [hex(got[y, x]) + " vs " + hex(ref[y, x]) for y, x in zip(ys[:3], xs[:3])])
from tests.test_oracle_ref_post_shaders import taa_inputs
hdr, depth, mv, hist, reproj = taa_inputs(rng, w, h)
ref_c, ref_h = oracle.taa_resolve(hdr, depth, mv, hist, reproj, 2)
oc = torch.zeros((h, w), dtype=torch.int32, device="cuda"); oh = harness.new_rgba16f(w, h)
harness.taa_resolve(harness.to_dev(hdr), harness.to_dev(depth), harness.to_dev(mv.reshape(h, w, 2)).view(torch.int32).reshape(h, w), harness.to_dev(hist), reproj, 2, oc, oh)
got_c, got_h = harness.to_host(oc, np.uint32), harness.to_host(oh, np.uint16)
dh = common.f16_ulp_diff(got_h, ref_h)
dc = np.max([np.abs(x - y) for x, y in zip(common.r11g11b10_codes(got_c), common.r11g11b10_codes(ref_c))], axis=0)
print("taa q2: history ulp max", dh.max(), "frac>0", (dh > 0).mean(), "frac>1", (dh > 1).mean(), "; colour max", dc.max(), "frac>0", (dc > 0).mean(), "frac>1", (dc > 1).mean())
ys, xs = np.nonzero((dh > 1).any(-1))
mvf = mv.view(np.float16).reshape(h, w, 2)
print("   bad examples", [(int(y), int(x), got_h[y, x].view(np.float16).tolist(), ref_h[y, x].view(np.float16).tolist(), mvf[y, x].tolist()) for y, x in zip(ys[:4], xs[:4])])
print("   bad with mv!=0:", (mvf[ys, xs] != 0).any(-1).mean() if len(ys) else None)
