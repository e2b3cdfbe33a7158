"""Colour render mode and surface export.

Colour: interp_tsdf_color (utils.cu:121-142) is dead code in the reference (its only call, viewer.cu:68, is
commented out), so no reference kernel produces a colour image; the checker launches the reference's own
device function (compiled verbatim into oracle/_ref) at the hit positions of our marcher -- which
test_gpu_raymarch.py pins to show_tsdf_kernel.  Surface export has no reference counterpart: checked against a
NumPy restatement of its definition on the downloaded planes."""
import numpy as np
import pytest

from tests.common import Scenario

pytestmark = pytest.mark.gpu


def fused_volume(sc):
    v = sc.make_volume()
    for fr in sc.frames:
        v.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
    return v


@pytest.mark.parametrize("angle", [0.1, 0.7])
def test_colour_render_matches_reference_interp_tsdf_color(angle):
    import torch
    from oracle import binding as ob
    from slam_maskrcnn_b200 import orbit_camera
    from tests.common import device_plane
    if not ob.ref_available(16):
        pytest.skip("oracle/_ref not built")
    sc = Scenario(dims=(64, 64, 64), bins=16, frames=6, yaw_step_deg=2.0)
    v = fused_volume(sc)
    s2w, c = orbit_camera(sc.Kinv, angle, float(sc.mean_depth))
    bgr, t, xyzt = v.raycast_color(s2w, c, want_t=True, want_xyzt=True)
    flags = v.ray_flags()
    hit = t > 0
    assert hit.mean() > 0.2 and (xyzt[..., 3] == t).all()
    assert (bgr[~hit] == 0).all(), "pixels without a surface must stay zero (viewer.cu:150)"
    # same rays, label mode: identical t (one marcher serves both)
    _, t2, _ = v.raycast(s2w, c, want_t=True, want_label=True)
    assert (t.view(np.uint32) == t2.view(np.uint32)).all()
    ok = hit & ((flags & 1) == 0)  # the reference reads out of bounds where a tap is clamped
    n = sc.W * sc.H
    xyz_d = torch.from_numpy(np.ascontiguousarray(xyzt[..., :3]).reshape(n, 3)).cuda()
    valid_d = torch.from_numpy(ok.reshape(n).astype(np.uint8)).cuda()
    out_d = torch.zeros(n * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    done = ob.ref_color_at(16, xyz_d.data_ptr(), valid_d.data_ptr(), n, sc.start, sc.voxel, sc.dims,
                           device_plane(v, "color").data_ptr(), out_d.data_ptr())
    if not done:
        pytest.skip("oracle/_ref predates the colour hook; rebuild it")
    ref = out_d.cpu().numpy().reshape(sc.H, sc.W, 3)
    same = (bgr == ref).all(axis=2) | ~ok
    assert same.all(), f"{int((~same).sum())} of {int(ok.sum())} colour pixels differ from the reference's interp_tsdf_color"
    assert bgr[ok].any()
    v.close()


def numpy_surface(sdf, wt, color, hist, start, voxel):
    """Definition of sfm_extract_surface on host planes: (xyz, bgr, label) sorted by position."""
    pts = []
    dims = sdf.shape
    obs = wt > 0
    idx = np.indices(dims).astype(np.float32)
    label = hist.argmax(axis=3).astype(np.uint8)  # first maximum == strict > scan in ascending order
    label[hist.max(axis=3) == 0] = 0
    for axis in range(3):
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[axis], sl1[axis] = slice(0, -1), slice(1, None)
        sl0, sl1 = tuple(sl0), tuple(sl1)
        s0, s1 = sdf[sl0], sdf[sl1]
        cross = obs[sl0] & obs[sl1] & ((s0 > 0) != (s1 > 0)) & (s0 != s1)
        with np.errstate(divide="ignore", invalid="ignore"):
            a = (s0 / (s0 - s1).astype(np.float32)).astype(np.float32)
        f = [idx[k][sl0].copy() for k in range(3)]
        f[axis] = f[axis] + a
        near0 = a <= 0.5
        xyz = np.stack([np.float32(f[k] * np.float32(voxel[k]) + np.float32(start[k])) for k in range(3)], axis=-1)[cross]
        bgr = np.where(near0[..., None], color[sl0], color[sl1])[cross]
        lab = np.where(near0, label[sl0], label[sl1])[cross]
        pts.append((xyz, bgr, lab))
    xyz = np.concatenate([p[0] for p in pts])
    bgr = np.concatenate([p[1] for p in pts])
    lab = np.concatenate([p[2] for p in pts])
    order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    return xyz[order], bgr[order], lab[order]


def test_surface_export_matches_its_definition(tmp_path):
    from slam_maskrcnn_b200 import write_ply
    sc = Scenario(dims=(48, 40, 56), bins=16, frames=6, yaw_step_deg=2.0)
    v = fused_volume(sc)
    xyz, bgr, lab = v.extract_surface()
    planes = {k: v.download(k) for k in ("sdf", "weight", "color", "hist")}
    rx, rb, rl = numpy_surface(planes["sdf"], planes["weight"], planes["color"], planes["hist"], sc.start, sc.voxel)
    assert len(xyz) == len(rx) > 1000
    # positions: fma on the device vs mul+add here -- equal to float rounding, so sort-stable comparison by tolerance
    assert np.allclose(xyz, rx, rtol=0, atol=2e-6 * max(1.0, float(np.abs(rx).max())))
    same_attr = (bgr == rb).all(axis=1) & (lab == rl)
    assert same_attr.mean() > 0.999, f"{int((~same_attr).sum())} points differ in colour/label (ties in the sort order only)"
    assert set(np.unique(lab)) <= set(range(16)) and (lab > 0).any()
    # every point lies inside the volume and within one voxel of an observed voxel
    lo, hi = np.asarray(sc.start, np.float32), np.asarray(sc.end, np.float32)
    assert (xyz >= lo - 1e-5).all() and (xyz <= hi + 1e-5).all()
    path = tmp_path / "surface.ply"
    write_ply(str(path), xyz, bgr, lab)
    head = path.read_bytes()[:300].decode("latin1")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(xyz)) and "property uchar label" in head
    assert path.stat().st_size == head.index("end_header\n") + len("end_header\n") + len(xyz) * 16
    v.close()
