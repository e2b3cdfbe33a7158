"""sfm_fuse_frame at the bench shape (512^3, 80 bins, 640x480): a few labelled frames with the duplicate-instance merge on;
prints wall-clock per frame.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--bins", type=int, default=80)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--instances", type=int, default=8)
    args = ap.parse_args()
    import bench
    from slam_maskrcnn_b200 import Volume, synth
    dims = tuple(args.dims)
    sc, K, Kinv, place, _ = bench.make_frames(1, dims, "tum")
    sc2 = synth.SynthScene(n_instances=args.instances, seed=1, yaw_step_deg=1.0, permute=True, hole_model="tum")
    frames = [sc2.frame(1 + i) for i in range(args.frames)]
    v = Volume(dims=dims, bins=args.bins, width=640, height=480, K=K, Kinv=Kinv)
    v.set_bounds(*place)
    ts = []
    for fr in frames:
        m = fr["mask"].copy()
        t0 = time.perf_counter()
        v.fuse_frame(fr["depth"], fr["color"], m, fr["extrinsic"])
        v.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print("ms per frame:", [round(t, 3) for t in ts], "num_objs", v.info().num_objs)
    v.close()


if __name__ == "__main__":
    main()
