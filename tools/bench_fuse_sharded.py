#!/usr/bin/env python3
"""Labelled fusion with duplicate-instance merging over z-slabs (BASELINE config 2 semantics on the
config 3 layout): `torchrun --nproc-per-node N tools/bench_fuse_sharded.py --gpus N`.  Rank 0 owns the
sequence; per frame: ncclBroadcast of the packed frame, exact sharded march from the incoming camera
(three NCCL MIN all-reduces), owner-side fold, NCCL SUM all-reduce of the integer overlap tables,
decision + relabel on every rank, integrate.  `--check` (small volumes) also fuses the sequence into a
whole volume on rank 0 with sfm_fuse_frame and asserts identical relabelled masks and num_objs."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dims", type=int, nargs=3, default=[256, 256, 256])
    ap.add_argument("--bins", type=int, default=32)
    ap.add_argument("--instances", type=int, default=8)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from slam_maskrcnn_b200 import Volume, synth
    from slam_maskrcnn_b200.slabs import SlabVolume, shard_halo, pack_frame, frame_nbytes, frame_offsets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = tuple(args.dims)
    sc = synth.SynthScene(n_instances=args.instances, seed=0, yaw_step_deg=2.0, permute=True)
    K = synth.intrinsic_matrix()
    Kinv = synth.intrinsic_inverse(K)
    f0 = sc.frame(0)
    md = synth.mean_depth(f0["depth"])
    place = synth.place_volume(f0["depth"], Kinv, md, dims)
    frames = [sc.frame(1 + i) for i in range(args.frames)]  # every rank builds the (deterministic) poses; rank 0's images are used
    sv = SlabVolume(dims, args.bins, rank, world, device=local, halo=shard_halo(place[2]) if world > 1 else 0, K=K, Kinv=Kinv)
    sv.set_bounds(*place)
    full = None
    if args.check and rank == 0:
        full = Volume(dims=dims, bins=args.bins, K=K, Kinv=Kinv, device=local)
        full.set_bounds(*place)
    nb = frame_nbytes()
    _, _, o_m, _ = frame_offsets()
    npx = 640 * 480
    buf = torch.empty(nb, dtype=torch.uint8, device="cuda")
    times, ok, objs = [], True, []
    for i, fr in enumerate(frames):
        if rank == 0:
            buf.copy_(torch.from_numpy(pack_frame(fr["depth"], fr["color"], fr["mask"], fr["extrinsic"])), non_blocking=False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sv.broadcast_frame(buf)
        sv.fuse_packed_sharded(buf, fr["extrinsic"])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
        objs.append(int(sv.vol.info().num_objs))
        if full is not None:
            m = fr["mask"].copy()
            full.fuse_frame(fr["depth"], fr["color"], m, fr["extrinsic"])
            got = buf[o_m:o_m + npx].cpu().numpy().reshape(m.shape)
            ok = ok and bool((got == m).all()) and full.info().num_objs == objs[-1]
    if rank == 0:
        out = {"what": "labelled fusion with duplicate-instance merge over z-slabs (broadcast + sharded march + fold + table all-reduce + decide + relabel + integrate)",
               "n_gpus": world, "dims": list(dims), "bins": args.bins, "frames": args.frames,
               "ms_per_frame_median": float(np.median(times[1:])), "ms_per_frame_first_merge": times[1] if len(times) > 1 else None,
               "num_objs": objs[-1], "instances_in_scene": args.instances}
        if args.check:
            out["matches_single_volume_fuse_frame"] = ok
        print(json.dumps(out))
        if args.check and not ok:
            sys.exit(1)
    sv.close()
    if full is not None:
        full.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
