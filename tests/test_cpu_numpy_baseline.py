"""The NumPy restatement of TSDF_Python/tsdf.py:78-120 (the CPU timing baseline) produces a sensible
fusion: same touched set as the C restatement of tsdf_kernel up to the documented semantic
differences (float64, trunc vs floor, `diff > -1` vs `diff > -miu`), SDF values close."""
import numpy as np

from tests.common import Scenario


def test_numpy_port_agrees_with_c_oracle_up_to_documented_differences():
    from oracle.tsdf_numpy import NumpyTSDF
    sc = Scenario(dims=(48, 48, 48), bins=0, frames=3)
    cv = sc.make_cpu_volume(bins=0)
    nt = NumpyTSDF(sc.intr, vol_dim=48)
    nt.set_bounds(sc.start, sc.end)
    for fr in sc.frames:
        cv.integrate(sc.K, fr["depth"], fr["color"], None, fr["extrinsic"], sc.W, sc.H)
        nt.integrate(fr["depth"], fr["color"], fr["extrinsic"].astype(np.float64))
    w_c, w_n = cv.wt, nt.tsdf_wt
    agree = (w_c == w_n).mean()
    assert agree > 0.995, f"weights agree on {agree:.4f} of the voxels"
    both = (w_c == w_n) & (w_c > 0)
    assert both.sum() > 1000
    assert np.abs(cv.sdf[both] - nt.tsdf_diff[both]).max() < 1e-3


def test_numpy_port_slab_mode_equals_whole_volume():
    from oracle.tsdf_numpy import NumpyTSDF
    sc = Scenario(dims=(32, 32, 32), bins=0, frames=2)
    a = NumpyTSDF(sc.intr, vol_dim=32)
    b = NumpyTSDF(sc.intr, vol_dim=32)
    a.set_bounds(sc.start, sc.end)
    b.set_bounds(sc.start, sc.end)
    for fr in sc.frames:
        E = fr["extrinsic"].astype(np.float64)
        a.integrate(fr["depth"], fr["color"], E)
        for x0 in range(0, 32, 8):
            b.integrate(fr["depth"], fr["color"], E, x_range=(x0, x0 + 8))
    assert (a.tsdf_wt == b.tsdf_wt).all() and (a.tsdf_diff == b.tsdf_diff).all() and (a.tsdf_color == b.tsdf_color).all()
