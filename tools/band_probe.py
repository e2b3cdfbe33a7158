"""Times the march of one rank's share of the image in the replicated-SDF ray-cast on one GPU: a 512^3 label-free volume
fused from the bench frames, 1280x960 orbit views; contiguous bands of 960/N rows (sfm_raycast_band_dev) against
interleaved 4-row tile rows (sfm_raycast_part_dev)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from slam_maskrcnn_b200 import Volume, orbit_camera, synth
    dims = (512, 512, 512)
    sc, K, Kinv, place, frames = bench.make_frames(12, dims, "tum")
    v = Volume(dims=dims, bins=0, width=640, height=480, K=K, Kinv=Kinv)
    v.set_bounds(*place)
    for fr in frames:
        v.integrate_raw(fr["depth"], fr["color"], fr["mask"], fr["extrinsic"])
    v.synchronize()
    w, h = 1280, 960
    K2 = np.array(K, np.float32).copy()
    K2[0, 0] *= 2; K2[1, 1] *= 2; K2[0, 2] *= 2; K2[1, 2] *= 2
    Kinv2 = synth.intrinsic_inverse(K2)
    md = synth.mean_depth(frames[0]["depth"])
    hits = torch.zeros(w * h * 4, dtype=torch.float32, device="cuda")
    cur = torch.cuda.current_stream().cuda_stream
    v.set_stream(cur)
    for n in (1, 2, 4, 8):
        rows = h // n
        res = []
        for band in range(n):
            angles = [0.05 + 0.37 * i for i in range(8)]
            for a in angles[:2]:
                s2w, c = orbit_camera(Kinv2, a, float(md))
                v.raycast_band_dev(s2w, c, w, h, band * rows, rows, hits.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for a in angles:
                s2w, c = orbit_camera(Kinv2, a, float(md))
                v.raycast_band_dev(s2w, c, w, h, band * rows, rows, hits.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / len(angles))
        print(f"N={n}: band march ms per view, per band: " + " ".join(f"{x:.3f}" for x in res) + f"  max {max(res):.3f}")
        res = []
        for part in range(n):
            angles = [0.05 + 0.37 * i for i in range(8)]
            for a in angles[:2]:
                s2w, c = orbit_camera(Kinv2, a, float(md))
                v.raycast_part_dev(s2w, c, w, h, part, n, hits.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for a in angles:
                s2w, c = orbit_camera(Kinv2, a, float(md))
                v.raycast_part_dev(s2w, c, w, h, part, n, hits.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / len(angles))
        print(f"N={n}: interleaved tile rows, per part:       " + " ".join(f"{x:.3f}" for x in res) + f"  max {max(res):.3f}")
    v.close()


if __name__ == "__main__":
    main()
