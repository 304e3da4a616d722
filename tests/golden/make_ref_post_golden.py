#!/usr/bin/env python
"""Regenerates tests/golden/refpost_*.npz: outputs of the REFERENCE's own post-processing shaders
(K7 threshold, K8 downsample +FEEDBACK, K9 upsample, K10 luminance, K11 tonemap, K12 FXAA, K13 TAA, HDR10 PQ encode),
executed on the CPU through the reference's vendored glslang + spirv-cross (`make -C oracle
ref-shaders`, oracle/ref_post_shim.cpp).  Needs /root/reference; the fixtures let machines without it
(the GPU box) check the oracle -- and through it the CUDA kernels -- against reference-derived vectors.

    python tests/golden/make_ref_post_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as oracle  # noqa: E402
from tests import common  # noqa: E402
from tests.test_oracle_ref_post_shaders import impls, pq_inputs, run_chain, taa_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    oracle.build()
    assert oracle.ref_post_kernels() is not None, "oracle/_ref post shaders were not built"
    _, ref = impls(oracle)
    w, h = 270, 135
    rng = np.random.default_rng(2701)
    hdr = common.random_hdr(rng, w, h)
    lum = np.array([0.3, 2.0 ** 0.3, 2.0 ** -0.3], np.float32)
    out = {"hdr": hdr, "lum_in": lum}
    hist = None
    for frame in range(2):
        b = run_chain(ref, hdr, lum, hist)
        for k, v in b.items():
            out[f"f{frame}_{k}"] = v
        lum, hist = b["lum"], b["d3"]
    np.savez_compressed(os.path.join(HERE, "refpost_chain_270x135.npz"), **out)

    w, h = 128, 80
    rng = np.random.default_rng(1280)
    hdr, depth, mv, hist, reproj = taa_inputs(rng, w, h)
    ldr = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    ldr[20:60, 30:90] = (ldr[20:60, 30:90] // 8) + 200
    ldr = np.ascontiguousarray(ldr).view(np.uint32).reshape(h, w)
    aa = {"hdr": hdr, "depth": depth, "mv": mv, "hist": hist, "reproj": reproj, "ldr": ldr,
          "fxaa_srgb": oracle.ref_fxaa(ldr, True), "fxaa_unorm": oracle.ref_fxaa(ldr, False)}
    for q in (0, 1, 2):
        aa[f"taa_q{q}_color"], aa[f"taa_q{q}_history"] = oracle.ref_taa_resolve(hdr, depth, mv, hist, reproj, q)
    aa["taa_first_color"], aa["taa_first_history"] = oracle.ref_taa_resolve(hdr, depth, mv, None, reproj, 2)
    np.savez_compressed(os.path.join(HERE, "refpost_aa_128x80.npz"), **aa)

    # HDR10 output encoding (pq10_encode.frag) with the matrix the reference's own math computes for BT.2020 primaries
    w, h = 96, 64
    rng = np.random.default_rng(2084)
    hdr, ui = pq_inputs(rng, w, h)
    m = oracle.ref_rec709_to_display_primaries(oracle.BT2020_PRIMARIES)
    pq = {"hdr": hdr, "ui": ui, "primaries": np.asarray(oracle.BT2020_PRIMARIES, np.float32), "primary_conversion": m,
          "pq10": oracle.ref_pq10_encode(hdr, ui, m, 500.0, 400.0, 1000.0)}
    np.savez_compressed(os.path.join(HERE, "refpost_pq10_96x64.npz"), **pq)
    print("written")


if __name__ == "__main__":
    main()
