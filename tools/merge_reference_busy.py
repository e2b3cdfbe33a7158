"""How the REFERENCE's duplicate-instance merge behaves on a busy scene: the verbatim back_proj_kernel +
TSDF::filter_overlaps + tsdf_kernel pipeline (oracle/_ref, tests/test_gpu_raymarch.run_reference_pipeline) next to
sfm_fuse_frame on the same sequence, frame by frame.  num_objs is unbounded in the reference (tsdf.cu:383); the run
stops when an id reaches the bin count, where the reference would write past its histogram (tsdf.cu:61).

    gpurun -- python tools/merge_reference_busy.py 79 80 24 > gpurun_out/merge_busy.log
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from oracle import binding as ob
    from tests.common import Scenario, backproj_camera
    ninst, bins, nframes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    yaw = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
    sc = Scenario(dims=(96, 96, 96), bins=bins, width=320, height=240, n_instances=ninst, frames=nframes, yaw_step_deg=yaw, permute=True)
    n = int(np.prod(sc.dims))
    pad = 2 * sc.dims[1] * sc.dims[2] + 64
    sdf = torch.full((n + 2 * pad,), float(sc.miu), dtype=torch.float32, device="cuda")
    wt = torch.zeros(n, dtype=torch.int32, device="cuda")
    col = torch.zeros(n * 3, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros((n + 2 * pad) * bins, dtype=torch.int32, device="cuda")
    sdf_p, cnt_p = sdf.data_ptr() + pad * 4, cnt.data_ptr() + pad * bins * 4
    ours = sc.make_volume()
    n_obs, num_objs = 0, 0
    print(f"# K = {ninst} instances in the scene, {bins} bins, {nframes} frames, yaw {yaw} deg/frame, labels permuted per frame + 5 % flips")
    for i, fr in enumerate(sc.frames):
        present = int((np.unique(fr["mask"]) > 0).sum())
        mask = fr["mask"].copy()
        if n_obs > 0:
            Rt, o = backproj_camera(fr["extrinsic"])
            probs = torch.zeros(sc.H * sc.W * bins, dtype=torch.float32, device="cuda")
            box = torch.zeros(sc.H * sc.W * bins, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            ob.ref_back_proj(bins, sc.Kinv, Rt, o, sc.start, sc.end, sc.voxel, sc.dims, sdf_p, cnt_p, sc.W, sc.H, probs.data_ptr(), box.data_ptr())
            mask, num_objs = ob.ref_filter_overlaps(probs.cpu().numpy(), mask, box.cpu().numpy(), bins, n_obs, num_objs)
        else:
            num_objs = int(mask.max()) + 1
        m2 = fr["mask"].copy()
        try:
            ours.fuse_frame(fr["depth"], fr["color"], m2, fr["extrinsic"])
            o_objs, o_same = ours.info().num_objs, bool((m2 == mask).all())
            note = ""
        except Exception as e:
            o_objs, o_same, note = None, None, f"  ours: {e}"
        print(f"frame {i + 1:3d}: labels present {present:3d}  reference num_objs {num_objs:4d}  ours {o_objs}  masks equal {o_same}{note}")
        if int(mask.max()) >= bins:
            print(f"# reference handed out id {int(mask.max())} >= {bins} bins: its next tsdf_kernel launch would index the histogram out of bounds; stopping")
            break
        d = torch.from_numpy(fr["depth"].view(np.int16)).cuda()
        c = torch.from_numpy(fr["color"]).cuda()
        m = torch.from_numpy(mask).cuda()
        torch.cuda.synchronize()
        ob.ref_integrate(bins, sdf_p, col.data_ptr(), cnt_p, wt.data_ptr(), sc.dims, sc.start, sc.voxel, float(sc.miu), sc.K,
                         d.data_ptr(), c.data_ptr(), m.data_ptr(), fr["extrinsic"], sc.W, sc.H)
        n_obs += 1
    ours.close()


if __name__ == "__main__":
    main()
