#!/usr/bin/env python
"""Regenerates tests/golden/refshader_*.npz: outputs of the REFERENCE's own clusterer compute
shaders (K1 spot transform, K2 cull setup, K3 binning, K4 Z range), executed on the CPU through the
reference's vendored glslang + spirv-cross (`make -C oracle ref-shaders`, oracle/ref_shader_shim.cpp).
Needs /root/reference; the fixtures let machines without it (the GPU box) still check the oracle
against reference-derived vectors.

    python tests/golden/make_ref_shader_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as oracle  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (aspect, lights, spot fraction, tile window (tx0, tx1, ty0, ty1) or None = whole 128x64 grid)
    "refshader_c1_16pt": (1.0, 16, 0.0, None),
    "refshader_300_25pct_spots": (16 / 9, 300, 0.25, (48, 80, 24, 40)),
}


def main():
    oracle.build()
    assert oracle.ref_kernels() is not None, "oracle/_ref reference shaders were not built"
    for name, (aspect, n, spots, window) in CASES.items():
        cam, lights, prep = common.build_lights_case(oracle, aspect, n, spots)
        r_spots = oracle.ref_spot_transform(cam, prep)
        r_cull = oracle.ref_cull_setup(cam, prep, r_spots)
        r_bitmask, r_fine = oracle.ref_binning(prep, r_cull, window)
        r_range = oracle.ref_z_range(prep)
        cam_arrays = {f"cam_{k}": np.array(list(getattr(cam, k)), np.float32) for k in
                      ("projection", "view", "view_projection", "inv_projection", "inv_view", "inv_view_projection", "camera_position", "camera_front")}
        cam_arrays["cam_z"] = np.array([cam.z_near, cam.z_far], np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), records=prep.records[:max(n, 1)].view(np.uint8), model=prep.model,
                            type_mask=prep.type_mask, z_ranges=prep.z_ranges, params=np.frombuffer(bytes(prep.params), np.uint8),
                            window=np.array(window if window else (0, 128, 0, 64), np.int32), n=np.array(n, np.int32),
                            ref_spots=r_spots, ref_cull=r_cull, ref_bitmask=r_bitmask, ref_bitmask_fine=r_fine, ref_range=r_range, **cam_arrays)
        print(name, "written")


if __name__ == "__main__":
    main()
