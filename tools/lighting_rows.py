#!/usr/bin/env python3
"""Time the lighting kernel alone over row bands of a bench frame (cluster built by the viewer).

    python tools/lighting_rows.py [--bands 1,2,4,8]

Answers whether the cost of a frame is additive over row bands (it decides how sharding scales).
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--bands", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--rows", default="", help="y0,y1: time only this band (for ncu)")
    args = ap.parse_args()
    os.environ["GRB_NO_ASYNC_POST"] = "1"
    os.environ["GRB_NO_ASYNC_CLUSTER"] = "1"
    import numpy as np
    import torch
    import bench
    from granite_b200 import capi, synth, viewer

    w, h, n_lights, aa, _ = bench.WORKLOADS[args.workload]
    scene = synth.make_scene(w, h)
    lights = synth.make_lights(n_lights, aspect=w / h)
    stream = torch.cuda.current_stream()
    v = viewer.Viewer(w, h, cuda_device=0, stream=stream.cuda_stream)
    v.set_camera(scene.projection, scene.view)
    v.set_directional(scene.dir_color, scene.dir_direction)
    v.set_lights(lights)
    v.bake()
    keep = [np.ascontiguousarray(a) for a in (scene.albedo, scene.normal, scene.pbr, scene.depth, scene.emissive)]
    v.render_frame(viewer.Viewer.host_gbuffer(*keep))
    v.sync()
    g = capi.GrbGBuffer()
    g.albedo, g.normal, g.pbr, g.depth, g.emissive = (v.image(n) for n in ("albedo", "normal", "pbr", "depth-transient", "emissive"))
    g.directional_color = (C.c_float * 3)(*scene.dir_color)
    g.directional_direction = (C.c_float * 3)(*scene.dir_direction)
    cam, _, _ = v.camera()
    params, bufs = v.cluster()
    hdr = v.image("HDR-main")
    lib = capi.lib()

    def run(rows):
        capi.check(lib.grb_deferred_lighting(C.byref(g), C.byref(cam), C.byref(params), C.byref(bufs), C.byref(hdr), capi.rows(rows),
                                             C.c_void_p(stream.cuda_stream)), "grb_deferred_lighting")

    def time_rows(rows):
        for _ in range(3):
            run(rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            run(rows)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps * 1000.0

    if args.rows:
        y0, y1 = (int(x) for x in args.rows.split(","))
        print(f"rows {y0}-{y1}: {time_rows((y0, y1)):.1f} us")
        v.close()
        return
    cost = viewer.estimate_band_cost(scene.projection, scene.view, lights.position, lights.color, w, h, depth=scene.depth, align=8)
    cost4 = v.measure_row_cost()
    print("measured cost: total %.1f M warp instructions, top 18 groups (72 rows) hold %.0f %%" % (cost4.sum() / 1e6, 100.0 * np.sort(cost4)[-18:].sum() / cost4.sum()))
    for n in [int(x) for x in args.bands.split(",")]:
        for label, bands in (("equal", viewer.band_partition(h, n, align=8)), ("modelled", viewer.band_partition_weighted(h, n, cost, align=8)),
                             ("measured", viewer.band_partition_measured(h, w, n, cost4, align=8))):
            if n == 1 and label != "equal":
                continue
            t = [time_rows(b) for b in bands]
            print(f"{n} {label:8s} bands: sum {sum(t):7.1f} us  max {max(t):7.1f} us   " + " ".join(f"{b[0]}-{b[1]}:{x:.0f}" for b, x in zip(bands, t)))
    v.close()


if __name__ == "__main__":
    main()
