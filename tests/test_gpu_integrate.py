"""K1 parity: the CUDA integrate path (through the C-ABI) against
  (1) the reference's own tsdf_kernel compiled verbatim (oracle/_ref) on the same GPU, and
  (2) the CPU C restatement (oracle/liboracle.so),
bit-exact on every plane: SDF (compared as raw bits), weight, colour, histogram."""
import numpy as np
import pytest

from tests.common import Scenario, bits

pytestmark = pytest.mark.gpu


def run_ours(sc, flags=0, labels=True, mask_key="gt"):
    v = sc.make_volume(flags=flags, bins=sc.bins if labels else 0)
    stats = []
    for fr in sc.frames:
        v.integrate_raw(fr["depth"], fr["color"], fr[mask_key] if labels else None, fr["extrinsic"])
        stats.append(v.frame_stats())
    out = {k: v.download(k) for k in (("sdf", "weight", "color", "hist") if labels else ("sdf", "weight", "color"))}
    v.close()
    return out, stats


def run_cpu_oracle(sc, labels=True, mask_key="gt"):
    cv = sc.make_cpu_volume(bins=sc.bins if labels else 0)
    stats = []
    for fr in sc.frames:
        stats.append(cv.integrate(sc.K, fr["depth"], fr["color"], fr[mask_key], fr["extrinsic"], sc.W, sc.H))
    return cv.planes(), stats


def run_reference_kernel(sc, mask_key="gt"):
    import torch
    from oracle import binding as ob
    n = int(np.prod(sc.dims))
    dev = "cuda"
    sdf = torch.full((n,), float(sc.miu), dtype=torch.float32, device=dev)
    wt = torch.zeros(n, dtype=torch.int32, device=dev)
    col = torch.zeros(n * 3, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(n * sc.bins, dtype=torch.int32, device=dev)
    for fr in sc.frames:
        d = torch.from_numpy(fr["depth"].view(np.int16)).to(dev)
        c = torch.from_numpy(fr["color"]).to(dev)
        m = torch.from_numpy(fr[mask_key]).to(dev)
        torch.cuda.synchronize()
        ob.ref_integrate(sc.bins, sdf.data_ptr(), col.data_ptr(), cnt.data_ptr(), wt.data_ptr(), sc.dims, sc.start,
                         sc.voxel, float(sc.miu), sc.K, d.data_ptr(), c.data_ptr(), m.data_ptr(), fr["extrinsic"], sc.W, sc.H)
    sh = sc.dims
    return {"sdf": sdf.cpu().numpy().reshape(sh), "weight": wt.cpu().numpy().reshape(sh),
            "color": col.cpu().numpy().reshape(sh + (3,)), "hist": cnt.cpu().numpy().view(np.uint32).reshape(sh + (sc.bins,))}


def assert_planes_equal(a, b, what):
    for k in a:
        if k == "sdf":
            same = bits(a[k]) == bits(b[k])
        else:
            same = a[k] == b[k]
        assert same.all(), f"{what}: plane {k} differs at {int((~same).sum())} of {same.size} entries"


@pytest.mark.parametrize("dims,bins", [((64, 64, 64), 16), ((48, 40, 36), 32), ((33, 31, 30), 16), ((40, 40, 29), 32)])
def test_integrate_matches_reference_kernel(dims, bins):
    from oracle import binding as ob
    if not ob.ref_available(bins):
        pytest.skip("oracle/_ref not built")
    sc = Scenario(dims=dims, bins=bins, frames=5)
    ours, stats = run_ours(sc)
    ref = run_reference_kernel(sc)
    assert_planes_equal(ours, ref, "vs reference tsdf_kernel")
    assert ours["weight"].sum() == sum(u for u, _ in stats)
    assert ours["hist"].sum() == sum(s for _, s in stats)
    assert ours["weight"].sum() > 0 and ours["hist"].sum() > 0


@pytest.mark.parametrize("dims", [(64, 64, 64), (33, 31, 30)])
def test_integrate_matches_cpu_oracle(dims):
    sc = Scenario(dims=dims, bins=16, frames=4)
    ours, stats = run_ours(sc)
    orc, ostats = run_cpu_oracle(sc)
    assert_planes_equal(ours, orc, "vs CPU oracle")
    assert stats == ostats


def test_cull_is_exact():
    """Brick culling must not change a single bit (it only skips voxels the reference rejects)."""
    from slam_maskrcnn_b200 import FLAG_NO_CULL
    sc = Scenario(dims=(96, 96, 96), bins=16, frames=6, yaw_step_deg=4.0)
    a, sa = run_ours(sc, flags=0)
    b, sb = run_ours(sc, flags=FLAG_NO_CULL)
    assert_planes_equal(a, b, "cull vs no-cull")
    assert sa == sb


def test_labels_off_mode():
    """bins == 0: a1 minus the histogram increment; SDF / weight / colour unchanged."""
    sc = Scenario(dims=(64, 64, 64), bins=16, frames=3)
    on, _ = run_ours(sc, labels=True)
    off, _ = run_ours(sc, labels=False)
    for k in ("sdf", "weight", "color"):
        assert (bits(on[k]) == bits(off[k])).all() if k == "sdf" else (on[k] == off[k]).all()


def test_permuted_noisy_labels_match_reference():
    from oracle import binding as ob
    if not ob.ref_available(16):
        pytest.skip("oracle/_ref not built")
    sc = Scenario(dims=(64, 64, 64), bins=16, frames=4, permute=True)
    ours, _ = run_ours(sc, mask_key="mask")
    ref = run_reference_kernel(sc, mask_key="mask")
    assert_planes_equal(ours, ref, "noisy labels vs reference tsdf_kernel")


def test_label_out_of_range_is_rejected():
    from slam_maskrcnn_b200 import SfmError
    sc = Scenario(dims=(32, 32, 32), bins=16, frames=1)
    v = sc.make_volume()
    fr = sc.frames[0]
    bad = fr["gt"].copy()
    bad[10, 10] = 16
    v.integrate_raw(fr["depth"], fr["color"], bad, fr["extrinsic"])
    with pytest.raises(SfmError):
        v.synchronize()
    v.close()


def test_full_size_frame_256_matches_reference_kernel():
    """BASELINE config 1 shape: 640x480 frames, 256^3, 16 bins (3 frames to keep it quick)."""
    from oracle import binding as ob
    if not ob.ref_available(16):
        pytest.skip("oracle/_ref not built")
    sc = Scenario(dims=(256, 256, 256), bins=16, width=640, height=480, n_instances=15, frames=3, yaw_step_deg=2.0)
    ours, stats = run_ours(sc)
    ref = run_reference_kernel(sc)
    assert_planes_equal(ours, ref, "256^3 vs reference tsdf_kernel")
    u = sum(u for u, _ in stats)
    assert 0.03 < u / (3 * 256 ** 3) < 0.4


def test_tma_staged_tile_grids_are_exact():
    """Tile grids staged into shared memory by cp.async.bulk (default) vs read through L1 (SFM_FLAG_NO_TMA)."""
    from slam_maskrcnn_b200 import FLAG_NO_TMA
    sc = Scenario(dims=(96, 96, 96), bins=16, frames=4, yaw_step_deg=4.0)
    a, sa = run_ours(sc, flags=0)
    b, sb = run_ours(sc, flags=FLAG_NO_TMA)
    assert_planes_equal(a, b, "TMA tiles vs L1 tiles")
    assert sa == sb


def test_pinhole_k_fast_path_is_exact():
    """K*c with the exact-zero terms dropped (default for a pinhole K) vs all nine terms (SFM_FLAG_GENERIC_K)."""
    from slam_maskrcnn_b200 import FLAG_GENERIC_K
    sc = Scenario(dims=(96, 96, 96), bins=16, frames=4, yaw_step_deg=4.0)
    a, sa = run_ours(sc, flags=0)
    b, sb = run_ours(sc, flags=FLAG_GENERIC_K)
    assert_planes_equal(a, b, "pinhole-K fast path vs generic K")
    assert sa == sb


@pytest.mark.parametrize("zl", [0, 1, 2, 3])
def test_brick_shapes_are_exact(zl, monkeypatch):
    """K1's brick shape ((32 >> zl) columns x (4 << zl) planes; thin z-slabs use flat bricks) never changes a bit."""
    sc = Scenario(dims=(64, 72, 64), bins=16, frames=4, yaw_step_deg=3.0)
    orc, ostats = run_cpu_oracle(sc)
    monkeypatch.setenv("SFM_ZL_LOG2", str(zl))
    ours, stats = run_ours(sc)
    assert_planes_equal(ours, orc, f"zl_log2={zl} vs CPU oracle")
    assert stats == ostats


@pytest.mark.parametrize("slab", [(20, 8), (12, 12), (8, 40), (0, 4), (28, 36)])
def test_thin_slabs_hold_identical_planes(slab):
    """A handle that stores only planes [z0, z0+nz) -- down to 4 planes, the brick shape follows nz -- holds
    the same bits as those planes of the whole volume."""
    sc = Scenario(dims=(64, 64, 64), bins=16, frames=4, yaw_step_deg=3.0)
    full, _ = run_ours(sc)
    z0, nz = slab
    v = sc.make_volume(slab=slab)
    for fr in sc.frames:
        v.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
    part = {k: v.download(k) for k in ("sdf", "weight", "color", "hist")}
    v.close()
    assert_planes_equal(part, {k: full[k][:, :, z0:z0 + nz] for k in part}, f"slab {slab} vs whole volume")
    assert part["weight"].sum() > 0


@pytest.mark.parametrize("blocks_per_sm", [1, 3])
def test_results_do_not_depend_on_scheduling(blocks_per_sm, monkeypatch):
    """Race check by construction (compute-sanitizer is closed on this pool, profiles/r2_sanitizer_*.log): the same
    sequence with a different number of resident K1b blocks per SM -- a different assignment of bricks to warps, a
    different interleaving of the surface-queue drains, and a preparation stream that runs further ahead -- and run
    twice in a row must leave identical planes.  Every voxel is written by exactly one lane per frame and the histogram
    bins by one reduction per voxel and frame, so any difference would be a race."""
    sc = Scenario(dims=(96, 96, 96), bins=16, frames=6, yaw_step_deg=4.0)
    a, sa = run_ours(sc)
    a2, sa2 = run_ours(sc)
    monkeypatch.setenv("SFM_K1B_BLOCKS_PER_SM", str(blocks_per_sm))
    b, sb = run_ours(sc)
    assert_planes_equal(a, a2, "run to run")
    assert_planes_equal(a, b, f"{blocks_per_sm} K1b blocks per SM vs default")
    assert sa == sa2 == sb
