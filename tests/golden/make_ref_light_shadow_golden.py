#!/usr/bin/env python
"""Regenerates tests/golden/reflight_shadows_160x96_300.npz: the lit HDR image of the REFERENCE's own clustering.frag
compiled with POSITIONAL_LIGHTS_SHADOW (+ directional.frag) run on the CPU (`make -C oracle ref-shaders`,
oracle/ref_light_shim.cpp KERNEL=7) for tests/test_oracle_ref_light_shadows.shadow_case(160, 96, 300 lights, 25 % spots,
32^2 synthetic shadow maps).  Needs /root/reference.

    python tests/golden/make_ref_light_shadow_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as oracle  # noqa: E402
from tests.test_oracle_ref_light_shadows import shadow_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    oracle.build()
    k = oracle.ref_light_kernels()
    assert k is not None and 7 in k, "oracle/_ref lighting shaders were not built"
    scene, cam, prep, clus, transforms, maps = shadow_case(oracle, 160, 96, 300, 0.25, 32)
    ref, _, c_rgb = oracle.ref_deferred_lighting(scene, cam, prep, clus, shadows=(transforms, maps, 32))
    np.savez_compressed(os.path.join(HERE, "reflight_shadows_160x96_300.npz"), ref_hdr=ref, clustered_rgb=c_rgb.astype(np.float32), depth=scene.depth,
                        transforms=transforms)
    print("written")


if __name__ == "__main__":
    main()
