"""Band-only variant of tools/band_probe.py (sfm_raycast_band_dev on a bins = 0 volume of D^3 voxels, 1280x960 views, the
whole image and the eight 1/8 bands): it only uses entry points that older builds of the library have, so that the same
script can time an exported older tree (git archive <commit> into build/oldsrc, built there) next to the current one.
PROBE_FLAGS: sfm_desc flags of the volume (32 = SFM_FLAG_DEBUG_ABLATE, then SFM_DEBUG_ABLATE=128 marches in raster order)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def main():
    import torch, bench
    from slam_maskrcnn_b200 import Volume, orbit_camera, synth
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dims = (D, D, D)
    sc, K, Kinv, place, frames = bench.make_frames(12, dims, "tum")
    v = Volume(dims=dims, bins=0, width=640, height=480, K=K, Kinv=Kinv, flags=int(os.environ.get("PROBE_FLAGS", "0")))
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
    v.set_stream(torch.cuda.current_stream().cuda_stream)
    for n in (1, 8):
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
        print(f"D={D} N={n}: band march ms per view: " + " ".join(f"{x:.3f}" for x in res) + f"  max {max(res):.3f}")
    v.close()
main()
