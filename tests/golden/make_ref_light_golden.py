#!/usr/bin/env python
"""Regenerates tests/golden/reflight_*.npz: the lit HDR image produced by the REFERENCE's own
directional.frag + clustering.frag run on the CPU (`make -C oracle ref-shaders`, oracle/ref_light_shim.cpp)
for the seeded synthetic scene tests/common.build_case(160, 96, 300 lights, 25 % spots).  Needs
/root/reference.

    python tests/golden/make_ref_light_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as oracle  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    oracle.build()
    assert oracle.ref_light_kernels() is not None, "oracle/_ref lighting shaders were not built"
    scene, cam, lights, prep = common.build_case(oracle, 160, 96, 300, 0.25)
    clus = oracle.cluster_build(cam, prep)
    ref, d_rgb, c_rgb = oracle.ref_deferred_lighting(scene, cam, prep, clus)
    np.savez_compressed(os.path.join(HERE, "reflight_160x96_300_25pct_spots.npz"), ref_hdr=ref, directional_rgb=d_rgb.astype(np.float32),
                        clustered_rgb=c_rgb.astype(np.float32), depth=scene.depth, albedo=scene.albedo)
    print("written")


if __name__ == "__main__":
    main()
