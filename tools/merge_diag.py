"""Diagnostic: how does the reference's duplicate-instance merge behave on the synthetic sequence?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slam_maskrcnn_b200 import Volume, synth

ninst, dims, bins, nframes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
yaw = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
sc = synth.SynthScene(n_instances=ninst, seed=0, yaw_step_deg=yaw, permute=True)
K = synth.intrinsic_matrix(); Kinv = synth.intrinsic_inverse(K)
f0 = sc.frame(0); md = synth.mean_depth(f0["depth"])
place = synth.place_volume(f0["depth"], Kinv, md, (dims,) * 3)
v = Volume(dims=(dims,) * 3, bins=bins, K=K, Kinv=Kinv)
v.set_bounds(*place)
for f in range(1, nframes + 1):
    fr = sc.frame(f)
    mask = fr["mask"].copy()
    present = np.unique(mask); present = present[present > 0]
    try:
        v.fuse_frame(fr["depth"], fr["color"], mask, fr["extrinsic"])
    except Exception as e:
        print("frame", f, "ERROR", e); break
    if f > 1:
        rep = v.last_merge()
        bp = np.array(rep.best_prob[:])[present]
        asg = np.array(rep.assign[:])[present]
        print(f"frame {f}: present {len(present)} num_objs {rep.num_objs} margin {rep.margin:.4f} best_prob min/med {bp.min():.3f}/{np.median(bp):.3f} below_thr {(bp <= 0.15).sum()}")
    else:
        print("frame 1: num_objs", v.info().num_objs)
