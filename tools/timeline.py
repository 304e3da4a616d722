#!/usr/bin/env python3
"""GPU timeline of a few resident frames: per pass begin/end relative to the first pass.

    python tools/timeline.py [--workload c3] [--frames 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/timeline.py

Shows what overlaps across the three streams (cluster build, lighting, post chain) and, with
several ranks, where a rank waits for its peers inside the sharded bloom pass.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--equal-bands", action="store_true")
    ap.add_argument("--bands", default="", help="explicit cuts, e.g. 792,944,1040 for 4 ranks")
    ap.add_argument("--show", default="", help="ranks to print (default all), e.g. 0,4")
    ap.add_argument("--compact", action="store_true", help="one line per frame: lighting start / duration and period, cluster / bloom / tonemap intervals")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import bench
    from granite_b200 import synth, viewer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h, n_lights, aa, _ = bench.WORKLOADS[args.workload]
    scene = synth.make_scene(w, h)
    lights = synth.make_lights(n_lights, aspect=w / h)
    stream = torch.cuda.Stream()
    v = viewer.Viewer(w, h, cuda_device=local_rank, timestamps=2, stream=stream.cuda_stream)
    v.set_camera(scene.projection, scene.view)
    v.set_directional(scene.dir_color, scene.dir_direction)
    v.set_lights(lights)
    bands = [(0, h)]
    if world > 1:
        if args.bands:
            cuts = [0] + [int(x) for x in args.bands.split(",")] + [h]
            bands = list(zip(cuts[:-1], cuts[1:]))
            assert len(bands) == world
        elif args.equal_bands:
            bands = viewer.band_partition(h, world)
        else:
            cost = viewer.estimate_band_cost(scene.projection, scene.view, lights.position, lights.color, w, h, depth=scene.depth, align=8)
            bands = viewer.band_partition_weighted(h, world, cost, align=8)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(viewer.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        v.init_collectives(bytes(uid.cpu().numpy().tobytes()), rank, world)
        v.set_row_shards(bands, rank)
    v.bake()
    keep = [np.ascontiguousarray(a) for a in (scene.albedo, scene.normal, scene.pbr, scene.depth, scene.emissive)]
    gb = viewer.Viewer.host_gbuffer(*keep)
    for _ in range(6):
        v.render_frame(gb)
    v.sync()
    v.collect_timeline()
    if world > 1:
        dist.barrier()
    for _ in range(args.frames):
        v.render_frame(None)
    v.sync()
    tl = v.collect_timeline()
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r != rank or (args.show and str(rank) not in args.show.split(",")):
            continue
        print(f"--- rank {rank} of {world}, rows {bands[rank] if world > 1 else (0, h)}")
        if args.compact:
            frames, cur = [], {}
            for name, b, e in tl:
                if name in cur:
                    frames.append(cur)
                    cur = {}
                cur[name] = (b * 1000, e * 1000)
            frames.append(cur)
            prev = None
            for i, f in enumerate(frames):
                L = f.get("lighting", (0, 0))
                line = f"frame {i:3d}: lighting {L[0]:8.1f} +{L[1] - L[0]:6.1f}  period {L[0] - prev if prev is not None else 0:6.1f} |"
                for k in ("clustering-bindless", "bloom-compute", "tonemap"):
                    if k in f:
                        line += f" {k.split('-')[0]} {f[k][0] - L[0]:+7.1f}..{f[k][1] - L[0]:+7.1f}"
                print(line)
                prev = L[0]
            continue
        for name, b, e in tl:
            print(f"{name:22s} {b * 1000:9.1f} -> {e * 1000:9.1f} us  ({(e - b) * 1000:7.1f})")
        sys.stdout.flush()
    v.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
