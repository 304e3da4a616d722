#!/usr/bin/env python3
"""Render a few resident frames of a bench workload and nothing else -- the command to wrap in ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none -s 96 -c 32 --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_frame.py --frames 10
    ncu --set full --clock-control none --import-source on -s 96 -c 16 \
        -o gpurun_out/frame_full python tools/ncu_frame.py --frames 10

One frame of the default workload is 16 launches (cluster 4, lighting 1, bloom 10, tonemap 1), so
`-s 96` skips six warm frames.  Streams are forced onto one queue (GRB_NO_ASYNC_*) so the launch
order in the report is the pass order.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--async-streams", action="store_true", help="keep the multi-stream schedule")
    args = ap.parse_args()
    if not args.async_streams:
        os.environ["GRB_NO_ASYNC_POST"] = "1"
        os.environ["GRB_NO_ASYNC_CLUSTER"] = "1"
    import numpy as np
    import bench
    from granite_b200 import synth, viewer

    w, h, n_lights, aa, _ = bench.WORKLOADS[args.workload]
    scene = synth.make_scene(w, h)
    lights = synth.make_lights(n_lights, aspect=w / h)
    v = viewer.Viewer(w, h, cuda_device=0, post_aa=viewer.AA_TAA_HIGH_PLUS_FXAA if aa == "taa+fxaa" else viewer.AA_NONE)
    v.set_camera(scene.projection, scene.view)
    v.set_directional(scene.dir_color, scene.dir_direction)
    v.set_lights(lights)
    v.bake()
    keep = [np.ascontiguousarray(a) for a in (scene.albedo, scene.normal, scene.pbr, scene.depth, scene.emissive)]
    gb = viewer.Viewer.host_gbuffer(*keep)
    v.render_frame(gb)
    for _ in range(args.frames - 1):
        v.render_frame(None)
    v.sync()
    print("rendered", args.frames, "frames of", args.workload)


if __name__ == "__main__":
    main()
