"""sfm_fuse_frame's failure path.  When the duplicate-instance merge would hand out a global id >= bins (the reference
writes past its histogram there, tsdf.cu:61,383) the call returns an error and must leave everything as it was: the
planes, the observation count, num_objs and the caller's mask -- relabel and update are gated on the device by the
decision's overflow flag, so a caller can drop or fix the frame and go on."""
import numpy as np
import pytest

from tests.common import Scenario, bits

pytestmark = pytest.mark.gpu


def test_failed_fuse_leaves_the_volume_untouched():
    from slam_maskrcnn_b200 import SfmError
    sc = Scenario(dims=(64, 64, 64), bins=4, n_instances=2, frames=8, yaw_step_deg=1.0)
    v = sc.make_volume()
    fr0 = sc.frames[0]
    m = fr0["gt"].copy()
    v.fuse_frame(fr0["depth"], fr0["color"], m, fr0["extrinsic"])   # first frame: no merge, num_objs = max label + 1
    v.set_num_objs(3)                                                # labels 0..2 in use, one bin left
    H, W = sc.H, sc.W
    # every later frame carries a made-up instance (label 3) on a different block of background surface: nothing in the
    # volume matches it, so the merge gives it a fresh id -- the first one fits (id 3), the second one cannot (id 4 >= bins)
    blocks = [(slice(8, 40), slice(4, 40)), (slice(60, 100), slice(110, 150)), (slice(70, 110), slice(50, 100)), (slice(10, 40), slice(100, 150))]
    failed_at = None
    for i, blk in enumerate(blocks):
        fr = sc.frames[1 + i]
        mask = fr["gt"].copy()
        sel = np.zeros((H, W), bool)
        sel[blk] = True
        sel &= (mask == 0) & (fr["depth"] > 0)
        assert sel.sum() > 200, "the made-up instance needs pixels on a surface"
        mask[sel] = 3
        before = {k: v.download(k) for k in ("sdf", "weight", "color", "hist")}
        info0 = v.info()
        given = mask.copy()
        try:
            v.fuse_frame(fr["depth"], fr["color"], mask, fr["extrinsic"])
        except SfmError as e:
            assert "bins" in str(e)
            failed_at = i
            after = {k: v.download(k) for k in before}
            for k in before:
                same = (bits(after[k]) == bits(before[k])) if k == "sdf" else (after[k] == before[k])
                assert same.all(), f"plane {k} changed although sfm_fuse_frame failed"
            info1 = v.info()
            assert info1.num_objs == info0.num_objs and info1.n_obs == info0.n_obs
            assert (mask == given).all(), "the caller's mask was relabelled although the call failed"
            # the volume is still usable: the same frame without the made-up instance goes through
            ok_mask = fr["gt"].copy()
            v.fuse_frame(fr["depth"], fr["color"], ok_mask, fr["extrinsic"])
            assert v.info().n_obs == info0.n_obs + 1 and v.info().num_objs == info0.num_objs
            assert int(v.download("weight").sum()) > int(before["weight"].sum())
            break
        assert v.info().num_objs <= sc.bins
    assert failed_at is not None and failed_at >= 1, "the merge never ran out of bins"
    v.close()
