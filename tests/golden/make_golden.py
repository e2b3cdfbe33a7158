#!/usr/bin/env python3
"""Generates the golden fixtures in this directory by running THE REFERENCE'S OWN CODE
(oracle/_ref: tsdf_kernel, back_proj_kernel, show_tsdf_kernel and TSDF::filter_overlaps compiled
verbatim from /root/reference/src/SfM_CUDA by oracle/build_ref.py) on small seeded inputs.

Needs a GPU for the three kernels:   gpurun -- python tests/golden/make_golden.py
Output: tests/golden/ref_small.npz (committed).  The CPU restatement (oracle/sfm_oracle.c) and the
CUDA path are both checked against it, so the oracle is pinned to real reference output rather
than to our reading of the source.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.common import Scenario, backproj_camera  # noqa: E402


def scenario():
    return Scenario(dims=(32, 32, 32), bins=16, width=80, height=60, n_instances=4, frames=4, seed=3, permute=True,
                    yaw_step_deg=3.0)


def main():
    import torch
    from oracle import binding as ob
    from slam_maskrcnn_b200 import synth
    sc = scenario()
    bins = sc.bins
    n = int(np.prod(sc.dims))
    pad = 2 * sc.dims[1] * sc.dims[2] + 64
    sdf = torch.full((n + 2 * pad,), float(sc.miu), dtype=torch.float32, device="cuda")
    wt = torch.zeros(n, dtype=torch.int32, device="cuda")
    col = torch.zeros((n + 2 * pad) * 3, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros((n + 2 * pad) * bins, dtype=torch.int32, device="cuda")
    sdf_p, cnt_p, col_p = sdf.data_ptr() + pad * 4, cnt.data_ptr() + pad * bins * 4, col.data_ptr() + pad * 3
    # (1) tsdf_kernel: 3 frames with ground-truth labels
    for fr in sc.frames[:3]:
        d = torch.from_numpy(fr["depth"].view(np.int16)).cuda()
        c = torch.from_numpy(fr["color"]).cuda()
        m = torch.from_numpy(fr["gt"]).cuda()
        torch.cuda.synchronize()
        ob.ref_integrate(bins, sdf_p, col_p, cnt_p, wt.data_ptr(), sc.dims, sc.start, sc.voxel, float(sc.miu), sc.K,
                         d.data_ptr(), c.data_ptr(), m.data_ptr(), fr["extrinsic"], sc.W, sc.H)
    out = {
        "sdf": sdf[pad:pad + n].cpu().numpy(), "weight": wt.cpu().numpy(),
        "color": col[pad * 3:(pad + n) * 3].cpu().numpy(),
        "hist": cnt[pad * bins:(pad + n) * bins].cpu().numpy().view(np.uint32),
    }
    # (2) back_proj_kernel from the 4th frame's pose
    fr = sc.frames[3]
    Rt, o = backproj_camera(fr["extrinsic"])
    probs = torch.zeros(sc.H * sc.W * bins, dtype=torch.float32, device="cuda")
    box = torch.zeros(sc.H * sc.W * bins, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ob.ref_back_proj(bins, sc.Kinv, Rt, o, sc.start, sc.end, sc.voxel, sc.dims, sdf_p, cnt_p, sc.W, sc.H,
                     probs.data_ptr(), box.data_ptr())
    out["probs"] = probs.cpu().numpy()
    out["box_mask"] = box.cpu().numpy()
    # (3) filter_overlaps (CPU, verbatim) on those probs and the 4th frame's permuted, noisy mask
    mask, num_objs = ob.ref_filter_overlaps(out["probs"], fr["mask"], out["box_mask"], bins, 3, int(sc.frames[0]["gt"].max()) + 1)
    out["merge_mask_in"] = fr["mask"]
    out["merge_mask_out"] = mask
    out["merge_num_objs"] = np.array([int(sc.frames[0]["gt"].max()) + 1, num_objs], np.int32)
    # (4) show_tsdf_kernel from an orbit camera (viewer.cu:140-146 matrices computed in float64 -> float32 here)
    angle, dist = 0.35, float(sc.mean_depth)
    rot = np.array([[np.cos(angle), 0, -np.sin(angle), dist * np.sin(angle)], [0, 1, 0, 0],
                    [np.sin(angle), 0, np.cos(angle), dist - dist * np.cos(angle)], [0, 0, 0, 1]], np.float32)
    s2w = (rot.astype(np.float64) @ sc.Kinv.astype(np.float64)).astype(np.float32)
    c3 = np.array([(dist + 0.5) * np.sin(angle), 0, (dist + 0.5) - (dist + 0.5) * np.cos(angle)], np.float32)
    from slam_maskrcnn_b200 import palette
    pal = palette(bins)  # the reference's own 16 colours (viewer.cu:93-109), what sfm_show uses
    img = torch.zeros(sc.H * sc.W * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ob.ref_show(bins, s2w, c3, sc.start, sc.end, sc.voxel, sc.dims, sdf_p, col_p, cnt_p, sc.W, sc.H, img.data_ptr(), pal)
    out["show_s2w"], out["show_c"], out["show_palette"] = s2w, c3, pal
    out["show_bgr"] = img.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for path in (os.path.join(HERE, "ref_small.npz"), os.path.join(ROOT, "gpurun_out", "ref_small.npz")):
        np.savez_compressed(path, **out)
    print("wrote ref_small.npz:", {k: v.shape for k, v in out.items()},
          "touched", int(out["weight"].sum()), "hist", int(out["hist"].sum()), "hits", int((out["probs"].reshape(-1, bins).sum(1) > 0).sum()),
          "num_objs", out["merge_num_objs"], "lit px", int((out["show_bgr"].reshape(-1, 3).sum(1) > 0).sum()))


if __name__ == "__main__":
    main()
