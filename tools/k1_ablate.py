"""K1 A/B tool: one process, one 512^3 x 80-bin volume, the bench's frames resident in HBM; runs the
integrate step under several SFM_DEBUG_ABLATE settings (one volume per setting) and prints the
mean CUDA-event time of K1 for each.  `SFM_B200_LIB=<path>` selects another build of the library.

    python tools/k1_ablate.py [--dims 512 512 512] [--bins 80] [--steps 60] [--ablate 0 8 16 24]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--bins", type=int, default=80)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--pool", type=int, default=12)
    ap.add_argument("--flags", type=int, nargs="*", default=[0])
    ap.add_argument("--ablate", type=int, nargs="*", default=[0])
    ap.add_argument("--slab", type=int, nargs=2, default=None)
    ap.add_argument("--in-order", action="store_true", help="frames become valid in stream order (no K1a overlap)")
    args = ap.parse_args()
    import torch
    import bench
    from slam_maskrcnn_b200 import Volume, _lib

    dims = tuple(args.dims)
    READY = Volume.READY_IN_ORDER if args.in_order else None
    sc, K, Kinv, place, frames = bench.make_frames(args.pool, dims, "tum")
    npx = 640 * 480
    packed = []
    for fr in frames:
        b = np.concatenate([fr["depth"].view(np.uint8).ravel(), fr["color"].ravel(), fr["gt"].ravel()])
        packed.append(torch.from_numpy(b).cuda())
    poses = [np.ascontiguousarray(fr["extrinsic"], dtype=np.float32) for fr in frames]
    out = {"lib": _lib.LIB_PATH, "dims": dims, "bins": args.bins, "runs": []}
    for flags in args.flags:
        for ab in args.ablate:
            # the library reads SFM_DEBUG_ABLATE once, at creation, and only under FLAG_DEBUG_ABLATE
            os.environ["SFM_DEBUG_ABLATE"] = str(ab)
            v = Volume(dims=dims, bins=args.bins, width=640, height=480, K=K, Kinv=Kinv, flags=flags | _lib.FLAG_DEBUG_ABLATE,
                       slab=tuple(args.slab) if args.slab else None)
            v.set_bounds(*place)
            for i in range(5):
                p = packed[i % len(packed)].data_ptr()
                v.integrate_dev(p, p + npx * 2, p + npx * 5, poses[i % len(packed)], ready=READY)
            v.synchronize()
            v.frame_stats()
            for i in range(args.steps):
                j = (i + 5) % len(packed)
                p = packed[j].data_ptr()
                v.integrate_dev(p, p + npx * 2, p + npx * 5, poses[j], ready=READY)
            v.synchronize()
            import time as _t
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            for i in range(args.steps):
                j = (i + 5) % len(packed)
                p = packed[j].data_ptr()
                v.integrate_dev(p, p + npx * 2, p + npx * 5, poses[j], ready=READY)
            v.synchronize()
            wall_ms = 1e3 * (_t.perf_counter() - t0) / args.steps
            ms = v.integrate_times(args.steps)
            ms_a, ms_b = v.integrate_times2(args.steps)
            U, S = v.frame_stats()
            r = {"flags": flags, "ablate": ab, "k1_ms_mean": float(np.mean(ms)), "k1_ms_median": float(np.median(ms)),
                 "k1_ms_min": float(np.min(ms)), "step_wall_ms": wall_ms, "k1a_ms": float(np.mean(ms_a)), "k1b_ms": float(np.mean(ms_b)), "U_per_step": U / args.steps, "S_per_step": S / args.steps}
            out["runs"].append(r)
            print(json.dumps(r), flush=True)
            v.close()
    os.environ.pop("SFM_DEBUG_ABLATE", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
