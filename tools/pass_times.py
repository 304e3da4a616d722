#!/usr/bin/env python3
"""Per-pass GPU times of a bench workload on one stream (CUDA events around every pass).

    python tools/pass_times.py [--workload c3] [--frames 40]

Used for A/B runs of kernel variants (e.g. GRB_LIGHTING_VARIANT=n); the scene is cached in /tmp so
that a loop over variants does not rebuild it.
"""
import argparse
import json
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    os.environ.setdefault("GRB_NO_ASYNC_POST", "1")
    os.environ.setdefault("GRB_NO_ASYNC_CLUSTER", "1")
    import numpy as np
    import bench
    from granite_b200 import synth, viewer

    w, h, n_lights, aa, _ = bench.WORKLOADS[args.workload]
    cache = f"/tmp/grb_scene_{args.workload}.pkl"
    if os.path.exists(cache):
        scene, lights = pickle.load(open(cache, "rb"))
    else:
        scene = synth.make_scene(w, h)
        lights = synth.make_lights(n_lights, aspect=w / h)
        pickle.dump((scene, lights), open(cache, "wb"))
    post = viewer.AA_TAA_HIGH_PLUS_FXAA if aa == "taa+fxaa" else viewer.AA_NONE
    v = viewer.Viewer(w, h, cuda_device=0, post_aa=post, timestamps=1)
    v.set_camera(scene.projection, scene.view)
    v.set_directional(scene.dir_color, scene.dir_direction)
    v.set_lights(lights)
    v.bake()
    keep = [np.ascontiguousarray(a) for a in (scene.albedo, scene.normal, scene.pbr, scene.depth, scene.emissive)]
    gb = viewer.Viewer.host_gbuffer(*keep)
    for _ in range(5):
        v.render_frame(gb)
    v.sync()
    v.collect_timings()
    for _ in range(args.frames):
        v.render_frame(None)
    v.sync()
    t = {k: round(ms / max(cnt, 1) * 1000.0, 1) for k, (ms, cnt) in v.collect_timings().items()}
    print(json.dumps({"tag": args.tag or os.environ.get("GRB_LIGHTING_VARIANT", ""), "pass_us": t}))


if __name__ == "__main__":
    main()
