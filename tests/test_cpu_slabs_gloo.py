"""Host-side logic of the multi-GPU path on CPU: slab partition, frame packing, and the two
collectives (frame broadcast, MIN composite of per-ray keys) over gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_the_volume():
    from slam_maskrcnn_b200.slabs import slab_range
    for dz in (128, 512, 1024, 100, 37):
        for world in (1, 2, 4, 8):
            if dz < world:
                continue
            z = 0
            for r in range(world):
                z0, nz = slab_range(r, world, dz)
                assert z0 == z and nz > 0
                if r < world - 1 and dz // world >= 4:
                    assert nz % 4 == 0
                z += nz
            assert z == dz


def test_frame_pack_roundtrip():
    from slam_maskrcnn_b200 import synth
    from slam_maskrcnn_b200.slabs import pack_frame, unpack_frame, frame_nbytes
    fr = synth.SynthScene(3).frame(2)
    buf = pack_frame(fr["depth"], fr["color"], fr["mask"], fr["extrinsic"])
    assert buf.nbytes == frame_nbytes() == 640 * 480 * 6 + 64
    d, c, m, p = unpack_frame(buf)
    assert (d == fr["depth"]).all() and (c == fr["color"]).all() and (m == fr["mask"]).all() and (p == fr["extrinsic"]).all()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from slam_maskrcnn_b200 import synth
    from slam_maskrcnn_b200.slabs import pack_frame, unpack_frame, composite_keys, slab_range, NO_HIT
    from tests.common import Scenario
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) frame broadcast: rank 0 owns the frame
        sc = Scenario(dims=(32, 32, 32), bins=8, frames=2)
        fr = sc.frames[1]
        buf = torch.from_numpy(pack_frame(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"]) if rank == 0
                               else np.zeros(sc.W * sc.H * 6 + 64, np.uint8))
        dist.broadcast(buf, src=0)
        d, c, m, p = unpack_frame(buf.numpy(), sc.W, sc.H)
        ok_bcast = bool((d == fr["depth"]).all() and (c == fr["color"]).all() and (m == fr["gt"]).all() and (p == fr["extrinsic"]).all())
        # (2) slab-sharded integration == whole-volume integration, plane by plane (CPU restatement per slab)
        z0, nz = slab_range(rank, world, sc.dims[2])
        whole = sc.make_cpu_volume()
        mine = sc.make_cpu_volume()
        for f in sc.frames:
            whole.integrate(sc.K, f["depth"], f["color"], f["gt"], f["extrinsic"], sc.W, sc.H)
            mine.integrate(sc.K, f["depth"], f["color"], f["gt"], f["extrinsic"], sc.W, sc.H, z_range=(z0, z0 + nz))
        w, s = whole.planes(), mine.planes()
        ok_slab = all((w[k][:, :, z0:z0 + nz] == s[k][:, :, z0:z0 + nz]).all() for k in ("weight", "color", "hist")) and \
            (w["sdf"][:, :, z0:z0 + nz].view(np.uint32) == s["sdf"][:, :, z0:z0 + nz].view(np.uint32)).all()
        untouched_elsewhere = (np.delete(s["weight"], np.s_[z0:z0 + nz], axis=2) == 0).all()
        # (3) min-composite of per-ray keys: each rank "hits" a different subset
        rng = np.random.default_rng(7)
        t = rng.uniform(0.5, 6.0, (world, 64)).astype(np.float32)
        lab = rng.integers(1, 8, (world, 64)).astype(np.int64)
        hit = rng.random((world, 64)) < 0.6
        keys_all = np.where(hit, (t.view(np.uint32).astype(np.int64) << 32) | lab, NO_HIT)
        mine_k = torch.from_numpy(keys_all[rank].copy())
        composite_keys(mine_k)
        ok_comp = bool((mine_k.numpy() == keys_all.min(0)).all())
        # the winner is the smallest t among the hitting ranks
        tt = np.where(hit, t, np.inf)
        win = tt.argmin(0)
        any_hit = hit.any(0)
        got_t = (mine_k.numpy() >> 32).astype(np.uint32).view(np.float32)
        ok_order = bool((got_t[any_hit] == t[win, np.arange(64)][any_hit]).all())
        q.put((rank, ok_bcast, ok_slab, bool(untouched_elsewhere), ok_comp, ok_order))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_broadcast_slabs_and_composite():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), f"rank {r[0]}: (bcast, slab, untouched, composite, order) = {r[1:]}"


def test_slab_planner_balances_a_front_loaded_profile():
    from slam_maskrcnn_b200.slabs import plan_slabs
    dz = 1024
    prof = np.where(np.arange(dz) < 560, 1.0, 0.0) * 0.8 / 560 + 0.2 / dz
    for world in (2, 4, 8):
        plan = plan_slabs(dz, world, prof)
        assert plan[0][0] == 0 and sum(n for _, n in plan) == dz
        assert all(plan[i][0] + plan[i][1] == plan[i + 1][0] for i in range(world - 1))
        assert all(n % 8 == 0 and 0 < n <= 3 * dz // world for _, n in plan)
        loads = [prof[z0:z0 + n].sum() * world for z0, n in plan]
        equal = [prof[i * dz // world:(i + 1) * dz // world].sum() * world for i in range(world)]
        assert max(loads) < 1.2 and max(loads) < max(equal)
    assert plan_slabs(512, 2) == [(0, 256), (256, 256)]


def test_slab_planner_splits_a_wall_over_thin_slabs():
    """A fronto-parallel wall puts most of the cost into ~10 planes: the plan must give them to thin slabs of
    their own instead of letting one rank own all of them plus free space."""
    from slam_maskrcnn_b200.slabs import plan_slabs, refine_profile
    dz, world = 1024, 8
    prof = np.full(dz, 0.03)
    prof[:680] += 1.0          # free space in front of the wall
    prof[672:688] += 40.0      # the wall's truncation band
    prof /= prof.sum()
    plan = plan_slabs(dz, world, prof)
    assert plan[0][0] == 0 and sum(n for _, n in plan) == dz
    assert all(n % 8 == 0 and 0 < n <= 3 * dz // world for _, n in plan)
    loads = np.array([prof[z0:z0 + n].sum() for z0, n in plan])
    wall_owners = [r for r, (z0, n) in enumerate(plan) if z0 < 688 and z0 + n > 672]
    assert len(wall_owners) >= 2, plan
    chunk_max = prof.reshape(-1, 8).sum(1).max()  # an 8-plane chunk cannot be split
    assert loads.max() <= chunk_max + 1e-9, (plan, loads)
    others = [l for r, l in enumerate(loads) if r not in wall_owners]
    assert max(others) < 1.3 / world, (plan, loads)
    # calibration: rescaling keeps the shape inside a slab, moves mass between slabs, stays normalised
    t = np.ones(world)
    t[wall_owners[0]] = 3.0
    prof2 = refine_profile(prof, plan, t)
    assert abs(prof2.sum() - 1.0) < 1e-12
    z0, n = plan[wall_owners[0]]
    assert prof2[z0:z0 + n].sum() == pytest.approx(3.0 / t.sum())
    seg, seg2 = prof[z0:z0 + n], prof2[z0:z0 + n]
    assert np.allclose(seg / seg.sum(), seg2 / seg2.sum())
    plan2 = plan_slabs(dz, world, prof2)
    assert plan2[wall_owners[0]][1] <= n  # the slab that measured slow does not grow


def _table_worker(rank, world, port, q):
    """The sharded merge's only SUM collective: int64 fixed-point tables + int32 counts, with FirstPix carried by
    rank 0 alone (zeros elsewhere) -- the layout sfm_shard_fold / SlabVolume.fuse_packed_sharded use."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = 16
        n64 = (2 * L * L + L) * 8
        n32 = (L * L + 4 * L) * 4
        rng = np.random.default_rng(11)
        parts64 = rng.integers(-2 ** 40, 2 ** 40, (world, n64 // 8), dtype=np.int64)
        parts32 = rng.integers(0, 2 ** 20, (world, n32 // 4), dtype=np.int64).astype(np.int32)
        first = np.full(L, -1, np.int32)            # 0xffffffff = label absent
        first[1:5] = [77, 3, 900, 12]
        parts32[:, -L:] = 0
        parts32[0, -L:] = first                      # only the counting rank carries FirstPix
        buf = torch.zeros(n64 + n32, dtype=torch.uint8)
        buf[:n64].view(torch.int64).copy_(torch.from_numpy(parts64[rank]))
        buf[n64:].view(torch.int32).copy_(torch.from_numpy(parts32[rank]))
        dist.all_reduce(buf[:n64].view(torch.int64), op=dist.ReduceOp.SUM)
        dist.all_reduce(buf[n64:].view(torch.int32), op=dist.ReduceOp.SUM)
        ok64 = bool((buf[:n64].view(torch.int64).numpy() == parts64.sum(0)).all())
        got32 = buf[n64:].view(torch.int32).numpy()
        ok32 = bool((got32[:-L] == parts32[:, :-L].sum(0, dtype=np.int32)).all())
        ok_first = bool((got32[-L:] == first).all())
        q.put((rank, ok64, ok32, ok_first))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_integer_table_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_table_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), f"rank {r[0]}: (int64 sum, int32 sum, first-pixel) = {r[1:]}"


def test_work_profile_z_cost_model_on_a_fake_volume():
    """work_profile_z with a stand-in volume (no GPU): the per-plane cost follows U and S at full z resolution,
    scaled from the coarse x/y grid to the fine one."""
    from slam_maskrcnn_b200 import slabs

    class FakeVolume:
        def __init__(self, dims):
            self.dims = dims
            self.n = 0

        def integrate_raw(self, depth, color, mask, extrinsic):
            assert not mask.any()  # the profile pass must use an all-zero label image (bin 0 counts S)
            self.n += 1

        def download(self, name):
            cx, cy, dz = self.dims
            if name == "weight":   # every frame touches planes [0, 60) completely
                w = np.zeros((cx, cy, dz), np.int32)
                w[:, :, :60] = self.n
                return w
            h = np.zeros((cx, cy, dz, 1), np.uint32)  # ... and updates the histogram in planes [50, 60)
            h[:, :, 50:60, 0] = self.n
            return h

        def close(self):
            pass

    frames = [{"depth": np.zeros((4, 4), np.uint16), "color": np.zeros((4, 4, 3), np.uint8), "gt": np.zeros((4, 4), np.uint8),
               "extrinsic": np.eye(4, dtype=np.float32)} for _ in range(3)]
    made = []

    def factory(pdims):
        made.append(pdims)
        return FakeVolume(pdims)

    prof, u, s = slabs.work_profile_z(factory, frames, (256, 512, 96), coarse_xy=64)
    assert made == [(64, 64, 96)]
    assert abs(prof.sum() - 1.0) < 1e-12 and len(prof) == 96
    assert np.allclose(u[:60], 256 * 512) and np.allclose(u[60:], 0)          # per frame, at the fine x/y resolution
    assert np.allclose(s[50:60], 256 * 512) and np.allclose(s[:50], 0)
    free, surf, empty = prof[10], prof[55], prof[80]
    assert surf / free == pytest.approx((slabs.COST_SURFACE + slabs.COST_VISIT) / (slabs.COST_FREE + slabs.COST_VISIT))
    assert empty / free == pytest.approx(slabs.COST_VISIT / (slabs.COST_FREE + slabs.COST_VISIT))
    plan = slabs.plan_slabs(96, 4, prof)
    assert sum(n for _, n in plan) == 96 and any(z0 >= 48 and n == 8 for z0, n in plan), plan


def _fuse_worker(rank, world, port, q):
    """SlabVolume.fuse_packed_sharded over gloo with a stand-in Volume (no GPU): the host side must issue the
    three MIN all-reduces between the march stages, let exactly rank 0 take the pixel counts, SUM the integer
    tables (int64 part and int32 part separately) and hand every rank the SAME reduced tables."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import ctypes
    import torch
    import torch.distributed as dist
    from slam_maskrcnn_b200 import slabs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, L = 8, 4, 4
        n = W * H
        n64, ntot = (2 * L * L + L) * 8, (2 * L * L + L) * 8 + (L * L + 4 * L) * 4

        def view(ptr, count, dtype):
            size = np.dtype(dtype).itemsize
            return np.frombuffer((ctypes.c_char * (count * size)).from_address(ptr), dtype=dtype)

        class Info:
            n_obs = 3

        class FakeVol:
            bins = L
            log = []

            def info(self):
                return Info()

            def shard_backproj_stage(self, stage, E, ev1, ev2, out):
                o = view(out, n, np.int64)
                if stage > 1:   # the previous stages arrive REDUCED: min over ranks of (1000*stage' + 10*rank + pixel)
                    assert (view(ev1, n, np.int64) == 1000 + np.arange(n)).all()
                if stage > 2:
                    assert (view(ev2, n, np.int64) == 2000 + np.arange(n)).all()
                o[:] = 1000 * stage + 10 * rank + np.arange(n)
                self.log.append(("stage", stage))

            def fold_table_bytes(self):
                return n64, ntot

            def shard_fold(self, d_mask, d_keys, do_counts, d_tables):
                assert (view(d_keys, n, np.int64) == 3000 + np.arange(n)).all()   # the composited keys
                assert bool(do_counts) == (rank == 0)
                t = view(d_tables, ntot, np.uint8)
                t[:n64].view(np.int64)[:] = (rank + 1) * 2 ** 40 + np.arange(n64 // 8)
                t[n64:].view(np.int32)[:] = (rank + 1) * 7
                self.log.append(("fold", bool(do_counts)))

            def shard_merge_finish(self, d_tables, d_mask):
                t = view(d_tables, ntot, np.uint8)
                tri = world * (world + 1) // 2
                ok64 = (t[:n64].view(np.int64) == tri * 2 ** 40 + world * np.arange(n64 // 8)).all()
                ok32 = (t[n64:].view(np.int32) == tri * 7).all()
                self.log.append(("finish", bool(ok64), bool(ok32)))
                m = view(d_mask, n, np.uint8)
                m[:] = 9   # "relabel"
                return np.arange(256, dtype=np.uint8), "report"

            def shard_first_frame(self, d_mask):
                self.log.append(("first",))

            def integrate_dev(self, d_depth, d_color, d_mask, E):
                self.log.append(("integrate", int(view(d_mask, n, np.uint8)[0])))

        sv = slabs.SlabVolume.__new__(slabs.SlabVolume)
        sv.rank, sv.world, sv.vol, sv.width, sv.height = rank, world, FakeVol(), W, H
        packed = torch.zeros(slabs.frame_nbytes(W, H), dtype=torch.uint8)
        lut, rep = sv.fuse_packed_sharded(packed, np.eye(4, dtype=np.float32))
        log = sv.vol.log
        ok = (log == [("stage", 1), ("stage", 2), ("stage", 3), ("fold", rank == 0), ("finish", True, True), ("integrate", 9)]
              and rep == "report" and (lut == np.arange(256)).all())
        q.put((rank, bool(ok), log))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_fuse_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_fuse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1], f"rank {r[0]}: {r[2]}"


def _replicated_raycast_worker(rank, world, port, q):
    """SlabVolume.raycast_replicated over gloo with stand-in volumes (no GPU): every rank marches ITS share of the image
    (4-row tile rows rank, rank + world, ... written densely), the shares are all-gathered part-major, the labelling step
    must find every pixel of the image at the documented place of the gathered buffer, and one MIN all-reduce must give
    every rank the same composite."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import ctypes
    import torch
    import torch.distributed as dist
    from slam_maskrcnn_b200 import slabs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H = 12, 10   # 3 tile rows for 2 ranks: the second share ends with a padding tile row
        no_hit = np.iinfo(np.int64).max

        def view(ptr, count, dtype):
            size = np.dtype(dtype).itemsize
            return np.frombuffer((ctypes.c_char * (count * size)).from_address(ptr), dtype=dtype)

        def part_rows(h, n):
            return 4 * (((h + 3) // 4 + n - 1) // n)

        class FakeReplica:
            def part_rows(self, h, n):
                return part_rows(h, n)

            def raycast_part_dev(self, s2w, c, w, h, part, n_parts, ptr):
                prow = part_rows(h, n_parts)
                buf = view(ptr, prow * w * 4, np.float32).reshape(prow, w, 4)
                buf[:] = 0
                for lr in range(prow // 4):
                    for k in range(4):
                        y = (part + lr * n_parts) * 4 + k
                        if y < h:
                            for x in range(w):
                                buf[lr * 4 + k, x] = (x, y, 0, 1 + y * w + x)

        class FakeSlab:
            class desc:
                device = 0

            def label_hits_parts_dev(self, ptr, w, h, n_parts, keys_ptr):
                prow = part_rows(h, n_parts)
                hits = view(ptr, n_parts * prow * w * 4, np.float32).reshape(n_parts, prow, w, 4)
                keys = view(keys_ptr, w * h, np.int64).reshape(h, w)
                for y in range(h):
                    gr = y // 4
                    for x in range(w):
                        e = hits[gr % n_parts, (gr // n_parts) * 4 + y % 4, x]
                        assert tuple(e) == (x, y, 0, 1 + y * w + x), f"pixel ({x},{y}) not where the share layout puts it: {e}"
                        owned = (x + y) % world == rank
                        keys[y, x] = (int(np.float32(e[3]).view(np.uint32)) << 32 | (rank + 1)) if owned else no_hit

        sv = slabs.SlabVolume.__new__(slabs.SlabVolume)
        sv.rank, sv.world, sv.vol, sv.replica, sv.width, sv.height = rank, world, FakeSlab(), FakeReplica(), W, H
        sv._host_device = torch.device("cpu")
        keys = sv.raycast_replicated(np.eye(4, dtype=np.float32), np.zeros(3, np.float32), W, H).numpy().reshape(H, W)
        yy, xx = np.mgrid[0:H, 0:W]
        t = (1 + yy * W + xx).astype(np.float32)
        want = (t.view(np.uint32).astype(np.int64) << 32) | ((xx + yy) % world + 1)
        q.put((rank, bool((keys == want).all()), int((keys != want).sum())))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_replicated_raycast_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_replicated_raycast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1], f"rank {r[0]}: {r[2]} pixels of the composite differ"


def test_slab_planner_properties_on_random_profiles():
    """plan_slabs on random per-plane cost profiles (spiky ones included): the slabs tile [0, dz) in order, every slab is a
    non-empty multiple of `align` planes no thicker than max_factor * dz / world, and the plan is never worse than equal
    thickness by the planner's own cost model (owned planes + the charged share of the high-side halo)."""
    from hypothesis import given, settings, strategies as st
    from slam_maskrcnn_b200 import slabs

    def plan_cost(plan, prof, align, halo_hi):
        c = np.asarray(prof, np.float64)
        c = c / c.sum()
        worst = 0.0
        for z0, n in plan:
            own = c[z0:z0 + n].sum()
            nxt = c[z0 + n:z0 + n + align].sum() * min(1.0, halo_hi / float(align))
            worst = max(worst, own + nxt)
        return worst

    @settings(max_examples=120, deadline=None)
    @given(st.integers(2, 8), st.integers(0, 2 ** 31 - 1), st.sampled_from([0, 4, 8]), st.sampled_from(["flat", "wall", "ramp", "noise"]))
    def check(world, seed, halo_hi, shape):
        align, dz = 8, 8 * 8 * world
        rng = np.random.default_rng(seed)
        prof = {"flat": np.ones(dz), "ramp": np.linspace(3.0, 0.05, dz), "noise": rng.random(dz) + 0.01,
                "wall": np.full(dz, 0.02)}[shape].copy()
        if shape == "wall":
            w0 = int(rng.integers(0, dz - 16))
            prof[w0:w0 + 12] += 5.0   # a fronto-parallel wall: a few very expensive planes
        plan = slabs.plan_slabs(dz, world, prof, align=align, halo_hi=halo_hi)
        assert len(plan) == world and plan[0][0] == 0
        z = 0
        for z0, n in plan:
            assert z0 == z and n > 0 and n % align == 0 and n <= int(3.0 * dz / world)
            z += n
        assert z == dz
        equal = [(r * (dz // world), dz // world) for r in range(world)]
        assert plan_cost(plan, prof, align, halo_hi) <= plan_cost(equal, prof, align, halo_hi) + 1e-9

    check()
