#!/usr/bin/env python3
"""Ray-cast bench (BASELINE config 4 shape): fuse a few labelled frames, then render orbit views at
1280x960 (intrinsics x2).  Single GPU: `python tools/bench_raycast.py`; sharded over N GPUs:
`torchrun --nproc-per-node N tools/bench_raycast.py --gpus N` (z-slabs with halos, three NCCL MIN
all-reduces per view).  `--check` also renders the views from a single volume on rank 0 and asserts that
the composited keys are identical (small volumes only)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dims", type=int, nargs=3, default=None)
    ap.add_argument("--bins", type=int, default=80)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--views", type=int, default=20)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=960)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from slam_maskrcnn_b200 import Volume, synth, orbit_camera
    from slam_maskrcnn_b200.slabs import SlabVolume, shard_halo, keys_to_int64

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = tuple(args.dims) if args.dims else {1: (512, 512, 512), 2: (512, 512, 1024), 4: (1024, 512, 1024), 8: (1024, 1024, 1024)}[world]
    sc = synth.SynthScene(n_instances=12, seed=0, yaw_step_deg=2.0, permute=False)
    K = synth.intrinsic_matrix()
    Kinv = synth.intrinsic_inverse(K)
    f0 = sc.frame(0)
    md = synth.mean_depth(f0["depth"])
    place = synth.place_volume(f0["depth"], Kinv, md, dims)
    frames = [sc.frame(1 + i) for i in range(args.frames)]
    # render camera: the viewer's orbit with intrinsics scaled to the output size (config 4: 1280x960 = x2)
    sx = args.width / 640.0
    Kr = synth.intrinsic_matrix(synth.FX * sx, synth.FY * sx, synth.CX * sx, synth.CY * sx)
    Kr_inv = synth.intrinsic_inverse(Kr)

    if world > 1:
        sv = SlabVolume(dims, args.bins, rank, world, device=local, halo=shard_halo(place[2]), K=K, Kinv=Kinv)
        vol = sv.vol
    else:
        vol = Volume(dims=dims, bins=args.bins, K=K, Kinv=Kinv, device=local)
        vol.set_stream(torch.cuda.current_stream().cuda_stream)
    vol.set_bounds(*place)
    for fr in frames:
        vol.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
    vol.synchronize()

    w, h = args.width, args.height
    keys = torch.empty(w * h, dtype=torch.int64, device="cuda")
    angles = [0.01 * (i + 1) * 15 for i in range(args.views)]  # kernel.cpp:104 steps by 0.01; sample every 15th view

    def render(a):
        s2w, c = orbit_camera(Kr_inv, a, float(md))
        if world > 1:
            return sv.raycast_sharded(s2w, c, w, h)
        vol.raycast_keys_dev(s2w, c, w, h, keys.data_ptr())
        return keys

    for a in angles[:3]:
        render(a)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    lit = 0
    for a in angles:
        k = render(a)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / len(angles)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    k64 = k if world > 1 else keys_to_int64(k)
    hits = int((k64 != np.iinfo(np.int64).max).sum())
    labelled = int(((k64 != np.iinfo(np.int64).max) & ((k64 & 0xff) > 0)).sum())

    ok = None
    if args.check:
        # single-volume reference on every rank (small volumes only), compared on rank 0
        ref_vol = Volume(dims=dims, bins=args.bins, K=K, Kinv=Kinv, device=local)
        ref_vol.set_stream(torch.cuda.current_stream().cuda_stream)
        ref_vol.set_bounds(*place)
        for fr in frames:
            ref_vol.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
        ok = True
        rk = torch.empty(w * h, dtype=torch.int64, device="cuda")
        for a in angles[:5]:
            s2w, c = orbit_camera(Kr_inv, a, float(md))
            ref_vol.raycast_keys_dev(s2w, c, w, h, rk.data_ptr())
            got = render(a)
            got = got if world > 1 else keys_to_int64(got)
            ok = ok and bool((keys_to_int64(rk) == got).all())
        ref_vol.close()

    if rank == 0:
        print(json.dumps({
            "metric": "rays/s", "value": w * h / (ms * 1e-3), "ms_per_view": ms, "n_gpus": world, "width": w, "height": h,
            "dims": list(dims), "bins": args.bins, "frames_fused": args.frames, "views": len(angles),
            "hit_rays_last_view": hits, "labelled_rays_last_view": labelled,
            "sharded": world > 1, "matches_single_volume": ok,
            "what": "march + shade (single GPU) / 3-stage exact sharded march with NCCL MIN all-reduces (multi GPU); keys stay on the device"}))
    vol.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
