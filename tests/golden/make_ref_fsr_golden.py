#!/usr/bin/env python
"""Regenerates tests/golden/reffsr_160x96_to_240x144.npz: the images of the REFERENCE's own upscale.frag and sharpen.frag
(FSR 1, 32-bit paths) run on the CPU (`make -C oracle ref-shaders`, oracle/ref_post_shim.cpp KERNEL 24 / 25 / 26) for the
seeded test image of tests/test_oracle_ref_smaa.smaa_test_image(160, 96, 7).  Needs /root/reference.

    python tests/golden/make_ref_fsr_golden.py
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
    k = oracle.ref_post_kernels()
    assert k is not None and 24 in k, "oracle/_ref post shaders were not built"
    img = smaa_test_image(160, 96, 7)
    up = oracle.ref_fsr_upscale(img, (240, 144))
    np.savez_compressed(os.path.join(HERE, "reffsr_160x96_to_240x144.npz"), color=img, upscaled_unorm=up,
                        upscaled_srgb=oracle.ref_fsr_upscale(img, (240, 144), target_srgb=True),
                        sharpened_srgb=oracle.ref_fsr_sharpen(up, 0.5, srgb=True), sharpened_unorm=oracle.ref_fsr_sharpen(up, 0.5, srgb=False))
    print("written")


if __name__ == "__main__":
    main()
