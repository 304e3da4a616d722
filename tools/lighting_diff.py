"""Diagnostic: where does the CUDA lighting pass differ from the oracle by more than one
B10G11R11 code?  (test infrastructure; prints a table, writes gpurun_out/lighting_diff.npz)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from granite_b200 import capi, harness, synth
from oracle import pyoracle as oracle
from tests import common


def main():
    w, h, n = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (3840, 2160, 4096)))
    oracle.build(ref=False)
    capi.lib()
    capi.init()
    scene, cam, lights, prep = common.build_case(oracle, w, h, n, 0.0)
    clus = oracle.cluster_build(cam, prep)
    ref, tile, zi, cnt = oracle.deferred_lighting(scene, cam, prep, clus, want_indices=True)
    dev = harness.ClusterDevice(prep.records, prep.model, prep.type_mask, prep.z_ranges, prep.params, prep.res)
    gcam = harness.camera_struct(cam)
    dev.build(gcam)
    gb = harness.GBufferDevice(scene)
    hdr = gb.emissive.clone()
    harness.deferred_lighting(gb, gcam, dev, hdr)
    got = harness.to_host(hdr, np.uint32)
    gc, rc = common.r11g11b10_codes(got), common.r11g11b10_codes(ref)
    diff = np.maximum.reduce([np.abs(a - b) for a, b in zip(gc, rc)])
    print("histogram of max code diff:", np.bincount(diff.reshape(-1))[:12])
    ys, xs = np.nonzero(diff > 1)
    print("pixels > 1 code:", len(ys))
    for y, x in list(zip(ys, xs))[:40]:
        print(f"({x},{y}) got={[int(c[y, x]) for c in gc]} ref={[int(c[y, x]) for c in rc]} lights={int(cnt[y, x])} depth={scene.depth[y, x]:.6g} "
              f"pbr={int(scene.pbr[y, x]):#06x} albedo={int(scene.albedo[y, x]):#010x} normal={int(scene.normal[y, x]):#010x} em={int(scene.emissive[y, x]):#010x}")
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/lighting_diff.npz", ys=ys, xs=xs, got=got[ys, xs], ref=ref[ys, xs], cnt=cnt[ys, xs])


if __name__ == "__main__":
    main()
