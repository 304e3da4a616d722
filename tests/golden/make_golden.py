#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (the reference holds no golden vectors for
this path -- SURVEY.md F4 -- so the oracle defines them; see oracle/oracle_math.h for what pins
the oracle itself).  Inputs are the seeded synthetic generators, stored alongside the outputs so
the GPU tests can replay them without the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from granite_b200 import synth  # noqa: E402
from oracle import pyoracle as oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def camera_arrays(cam):
    return {f"cam_{k}": np.array(list(getattr(cam, k)), np.float32) for k in
            ("projection", "view", "view_projection", "inv_projection", "inv_view", "inv_view_projection", "camera_position", "camera_front")} | {
        "cam_z": np.array([cam.z_near, cam.z_far], np.float32)}


def frame_case(name, w, h, n_lights, spots, frames=2):
    scene = synth.make_scene(w, h)
    cam = oracle.camera_setup(scene.projection, scene.view)
    lights = synth.make_lights(n_lights, spot_fraction=spots, aspect=w / h)
    prep = oracle.prepare_lights(cam, lights)
    clus = oracle.cluster_build(cam, prep)
    hdr, tile, zi, cnt = oracle.deferred_lighting(scene, cam, prep, clus, want_indices=True)
    out = dict(albedo=scene.albedo, normal=scene.normal, pbr=scene.pbr, depth=scene.depth, emissive=scene.emissive,
               dir_color=np.array(scene.dir_color, np.float32), dir_direction=np.array(scene.dir_direction, np.float32),
               records=prep.records[:max(n_lights, 1)].view(np.uint8), model=prep.model, type_mask=prep.type_mask, z_ranges=prep.z_ranges,
               params=np.frombuffer(bytes(prep.params), np.uint8),
               spots=clus.spots, cull=clus.cull, bitmask=clus.bitmask, cluster_range=clus.range,
               hdr=hdr, tile_index=tile, z_index=zi, light_count=cnt, **camera_arrays(cam))
    lum, d3 = np.zeros(3, np.float32), None
    for i in range(frames):
        f = oracle.hdr_chain(hdr, lum, d3)
        lum, d3 = f.lum, f.d3
        for k in ("t", "d0", "d1", "d2", "d3", "u2", "u1", "u0", "lum", "ldr"):
            out[f"f{i}_{k}"] = getattr(f, k)
    out["fxaa_srgb"] = oracle.fxaa(f.ldr, True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k in ("bitmask", "hdr", "f1_ldr")}, "lights/pixel max", int(cnt.max()))


def taa_case(name, w, h):
    rng = np.random.default_rng(99)
    rgb = (rng.random((h, w, 3)) ** 3 * 3.0).astype(np.float32)
    hdr = synth.pack_r11g11b10(rgb)
    depth = rng.uniform(0.0005, 0.03, size=(h, w)).astype(np.float32)
    depth[rng.random((h, w)) < 0.1] = 0.0
    mv = np.zeros((h, w, 2), np.float16)
    m = rng.random((h, w)) < 0.1
    mv[m] = (rng.uniform(-2.0, 2.0, size=(int(m.sum()), 2)) / np.array([w, h])).astype(np.float16)
    reproj = np.array([[0.5, 0, 0, 0], [0, 0.5, 0, 0], [0.3, -0.2, 1, 0], [0.5 + 0.4 / w, 0.5 - 0.3 / h, 0, 1]], np.float32)
    out = dict(hdr=hdr, depth=depth, mv=mv.view(np.uint16), reproj=reproj)
    for q in (0, 1, 2):
        c0, h0 = oracle.taa_resolve(hdr, depth, mv.view(np.uint16), None, reproj, q)
        c1, h1 = oracle.taa_resolve(hdr, depth, mv.view(np.uint16), h0, reproj, q)
        out[f"q{q}_color0"], out[f"q{q}_hist0"], out[f"q{q}_color1"], out[f"q{q}_hist1"] = c0, h0, c1, h1
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written")


if __name__ == "__main__":
    oracle.build(ref=False)
    frame_case("frame_96x64_40lights", 96, 64, 40, 0.25)
    frame_case("frame_c1_256x256_16lights", 256, 256, 16, 0.0, frames=1)
    taa_case("taa_80x48", 80, 48)
