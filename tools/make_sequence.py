#!/usr/bin/env python3
"""Writes a synthetic sequence in the on-disk layout the reference's driver reads (kernel.cpp:41-58):
<out>/rgb/<ts>.png (8-bit colour), <out>/depth/<ts>.png (16-bit, metres*5000), <out>/mask/<ts>.png.png
(8-bit Mask R-CNN label image, Mask_RCNN/dmask.py:47-58) and <out>/groundtruth.txt (TUM format)."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_maskrcnn_b200 import synth  # noqa: E402


def main(out, n_frames=20, n_instances=6, yaw=0.5):
    for d in ("rgb", "depth", "mask"):
        os.makedirs(os.path.join(out, d), exist_ok=True)
    sc = synth.SynthScene(n_instances=n_instances, yaw_step_deg=yaw, permute=True)
    with open(os.path.join(out, "groundtruth.txt"), "w") as gt:
        gt.write("# ground truth trajectory\n# timestamp tx ty tz qx qy qz qw\n")
        for f in range(n_frames):
            fr = sc.frame(f)
            p = sc.tum_pose(f, t0=1311868164.05, dt=0.25)
            ts = f"{p[0]:.6f}"
            gt.write(ts + " " + " ".join(f"{v:.9f}" for v in p[1:]) + "\n")
            cv2.imwrite(os.path.join(out, "depth", ts + ".png"), fr["depth"])
            cv2.imwrite(os.path.join(out, "rgb", ts + ".png"), fr["color"])  # cv2 writes BGR arrays as the reference reads them
            cv2.imwrite(os.path.join(out, "mask", ts + ".png.png"), fr["mask"])
    print("wrote", n_frames, "frames to", out)


if __name__ == "__main__":
    main(sys.argv[1], *(int(a) for a in sys.argv[2:4]))
