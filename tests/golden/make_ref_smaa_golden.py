#!/usr/bin/env python
"""Regenerates tests/golden/refsmaa_160x96.npz: the REFERENCE's SMAA shaders (edge detection, blending weights,
neighbourhood blending, presets Low .. Ultra) executed on the CPU (`make -C oracle ref-shaders`,
oracle/ref_post_shim.cpp) on one test image, together with the lookup textures they sampled
(assets/textures/smaa/{area,search}.gtx payloads: inputs of the passes, needed wherever the fixture is replayed).
Needs /root/reference.

    python tests/golden/make_ref_smaa_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as oracle  # noqa: E402
from tests.test_oracle_ref_smaa import smaa_test_image  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    oracle.build()
    assert oracle.ref_post_kernels() is not None, "oracle/_ref post shaders were not built"
    area, search = oracle.smaa_luts()
    img = smaa_test_image(160, 96, 7)
    out = {"color": img, "area": area, "search": search}
    for q in range(4):
        e = oracle.ref_smaa_edge(img, q)
        w = oracle.ref_smaa_weights(e, area, search, q)
        out[f"q{q}_edges"], out[f"q{q}_weights"], out[f"q{q}_out"] = e, w, oracle.ref_smaa_blend(img, w, q)
    np.savez_compressed(os.path.join(HERE, "refsmaa_160x96.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
