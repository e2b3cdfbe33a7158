"""z-slab sharding of one volume over the GPUs of a box: one process per GPU (torch.distributed).

The reference is single-GPU; this is the scale-out of its path.  Integration
is per-voxel independent (tsdf.cu:55-68 touches only its own voxel), so rank r simply owns the
global z planes [z0, z0+nz) and integrates every frame into them; the only exchange is the frame
itself, which rank `src` owns and broadcasts (ncclBroadcast over NVLink; gloo in the CPU tests).
Voxel positions are computed from GLOBAL z indices (fma(z_global, voxel.z, start.z)), so a slab
holds bit-identical values to the same planes of a single-GPU volume.

Ray-casting a sharded volume composites per-ray first-hit keys with a MIN all-reduce
(`composite_keys`): key = float_bits(t_hit) << 32 | label; t > 0 so the bit pattern orders like t.
"""
import numpy as np

FRAME_W, FRAME_H = 640, 480


def slab_range(rank, world, dz, align=4):
    """Global z planes [z0, z0+nz) owned by `rank`: as even as possible, boundaries on multiples of
    `align` (the 128-bit plane path needs nz % 4 == 0), the last rank takes the remainder."""
    assert 0 <= rank < world and dz >= world
    per = (dz // world) // align * align
    if per == 0:
        per, align = dz // world, 1
    z0 = rank * per
    nz = per if rank < world - 1 else dz - z0
    return z0, nz


def work_profile(volume_factory, frames, dz, coarse=128, fixed_share=0.2):
    """Relative integration cost of every global z plane, estimated on the GPU: the given frames are
    integrated into a coarse volume of the same physical cube (`volume_factory((coarse,)*3)` must return
    a placed, labels-off Volume), the weight plane gives the touched voxels per z, and a constant term
    stands for the per-brick classification every plane pays (about 20 % of the kernel at 512^3).
    Under brick culling the near planes (free space in front of the surfaces) carry almost all of the
    updates and the planes behind the surfaces none, so equal-thickness z-slabs are badly balanced."""
    vol = volume_factory((coarse, coarse, coarse))
    for fr in frames:
        vol.integrate_raw(fr["depth"], fr["color"], None, fr["extrinsic"])
    w = vol.download("weight").astype(np.float64).sum(axis=(0, 1))  # touched voxel-frames per coarse z plane
    vol.close()
    w = w / max(w.sum(), 1.0)
    cost = fixed_share / coarse + (1.0 - fixed_share) * w
    # resample to dz planes (piecewise constant), keep the total
    idx = np.minimum((np.arange(dz) * coarse) // dz, coarse - 1)
    prof = cost[idx]
    return prof / prof.sum()


# Relative cost of K1 per voxel of each kind, measured on B200 (profiles/README.md): a voxel that only
# gets its weight bumped (free space), a near-surface voxel (colour + histogram read-modify-write: bound
# by random 32-byte DRAM sector traffic when a rank owns little else), and the classification every voxel
# of a plane pays whether it is touched or not.
COST_FREE, COST_SURFACE, COST_VISIT = 1.0, 11.0, 0.03


def work_profile_z(volume_factory, frames, dims, coarse_xy=128):
    """Per-plane integration cost at FULL z resolution, estimated on the GPU.  The truncation band is
    5 voxels of the fine volume thick, so a fronto-parallel wall concentrates all its colour/histogram
    updates in ~10 fine planes: a profile taken on a coarser z grid smears exactly the feature the slab
    plan has to resolve.  `volume_factory((cx, cy, dz))` must return a placed Volume with bins = 1 whose
    x/y voxels are coarser but whose z voxel size and truncation distance are the fine volume's.  The
    frames are integrated with an all-zero label image: the weight plane then counts the touched voxels
    per plane (U) and histogram bin 0 the near-surface voxels (S)."""
    dx, dy, dz = dims
    cx, cy = min(coarse_xy, dx), min(coarse_xy, dy)
    vol = volume_factory((cx, cy, dz))
    zero = np.zeros_like(frames[0]["gt"])
    for fr in frames:
        vol.integrate_raw(fr["depth"], fr["color"], zero, fr["extrinsic"])
    u = vol.download("weight").astype(np.float64).sum(axis=(0, 1))
    s_ = vol.download("hist").astype(np.float64).sum(axis=(0, 1, 3))
    vol.close()
    scale = (dx * dy) / float(cx * cy) / max(len(frames), 1)
    u, s_ = u * scale, s_ * scale
    cost = COST_FREE * (u - s_) + COST_SURFACE * s_ + COST_VISIT * dx * dy
    return cost / cost.sum(), u, s_


def plan_slabs(dz, world, profile=None, align=8, max_factor=3.0, halo_hi=0):
    """Contiguous z-slabs [(z0, nz)] for `world` ranks.  Without a profile: equal thickness.  With a
    per-plane cost profile: the partition minimising the largest slab cost subject to every slab being
    a multiple of `align` planes and at most max_factor * dz / world planes thick (memory: the
    histogram of a slab must fit one GPU).  Greedy sweep inside a bisection on the cost bound.
    `halo_hi`: planes every slab stores (and integrates) beyond its owned range on the high-z side; their cost is
    charged to the slab (a fraction halo_hi / align of the next chunk), which matters for the thin slabs a
    fronto-parallel wall ends up in."""
    if profile is None or world == 1:
        return [slab_range(r, world, dz, align=min(align, 4)) for r in range(world)]
    assert len(profile) == dz and dz % align == 0 and dz // align >= world
    cost = np.asarray(profile, np.float64).reshape(dz // align, align).sum(1)
    cost = cost / cost.sum()
    nchunks = len(cost)
    hfrac = min(1.0, halo_hi / float(align))
    halo_cost = np.append(cost[1:], 0.0) * hfrac  # halo_cost[i] = what a slab ENDING with chunk i pays for its halo
    max_c = max(1, int(max_factor * dz / world) // align)
    assert max_c * world >= nchunks, "max_factor too small to cover the volume"

    def sweep(bound):
        cuts, i = [0], 0
        for r in range(world):
            acc, n = 0.0, 0
            remaining_slabs = world - r - 1
            # leave >= 1 chunk per later slab, and no more than the later slabs can hold
            while i < nchunks and n < max_c and nchunks - i > remaining_slabs:
                if n > 0 and acc + cost[i] + halo_cost[i] > bound and (nchunks - i) <= remaining_slabs * max_c:
                    break
                acc += cost[i]
                i += 1
                n += 1
            cuts.append(i)
        return cuts if i == nchunks else None

    def worst(cuts):
        return max(cost[cuts[r]:cuts[r + 1]].sum() + halo_cost[cuts[r + 1] - 1] for r in range(world))

    lo, hi = 1.0 / world, 1.0
    best = sweep(hi)
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        c = sweep(mid)
        if c is not None and worst(c) <= mid + 1e-12:
            best, hi = c, mid
        else:
            lo = mid
    cap = worst(best) + 1e-12

    # The min-max partition is not unique (one unsplittable chunk can set the bound and leave the greedy
    # sweep free to overfill the other slabs up to it).  Second pass: every slab aims at an even share of
    # what is left, never exceeds the bound, and always leaves a remainder the later slabs can hold.
    def feasible(i, slabs_left):
        for _ in range(slabs_left):
            acc, n = 0.0, 0
            while i < nchunks and n < max_c and (n == 0 or acc + cost[i] + halo_cost[i] <= cap):
                acc += cost[i]
                i += 1
                n += 1
        return i == nchunks

    cuts, i = [0], 0
    for r in range(world):
        left = world - r - 1
        desired = cost[i:].sum() / (world - r)
        acc, n = 0.0, 0
        while i + n < nchunks - left and n < max_c:
            c = cost[i + n]
            if n > 0 and acc + c + halo_cost[i + n] > cap:
                break
            if n > 0 and acc + 0.5 * c > desired and feasible(i + n, left):
                break
            acc += c
            n += 1
        i += n
        cuts.append(i)
    if i == nchunks and all(b > a for a, b in zip(cuts, cuts[1:])) and worst(cuts) <= cap:
        best = cuts
    return [(best[r] * align, (best[r + 1] - best[r]) * align) for r in range(world)]


def refine_profile(profile, plan, measured_ms):
    """Calibration step: rescale the per-plane profile inside every slab so that the slab's share of the
    profile equals its share of the measured kernel time (the SHAPE inside a slab is kept -- a flat
    per-slab density would smear a wall that sits at one end of a slab)."""
    prof = np.asarray(profile, np.float64).copy()
    t = np.asarray(measured_ms, np.float64)
    for (z0, n), ti in zip(plan, t):
        seg = prof[z0:z0 + n]
        tot = seg.sum()
        prof[z0:z0 + n] = (seg / tot if tot > 0 else np.full(n, 1.0 / n)) * ti
    return prof / prof.sum()


def frame_nbytes(width=FRAME_W, height=FRAME_H):
    return width * height * 6 + 64


def pack_frame(depth, color, mask, extrinsic2init, out=None):
    """[depth u16 | colour u8x3 | mask u8 | extrinsic2init f32x16] as one byte buffer (one broadcast)."""
    h, w = depth.shape
    n = w * h
    buf = np.empty(frame_nbytes(w, h), np.uint8) if out is None else out
    buf[:2 * n] = np.ascontiguousarray(depth, np.uint16).reshape(-1).view(np.uint8)
    buf[2 * n:5 * n] = np.ascontiguousarray(color, np.uint8).reshape(-1)
    buf[5 * n:6 * n] = np.ascontiguousarray(mask, np.uint8).reshape(-1)
    buf[6 * n:] = np.ascontiguousarray(extrinsic2init, np.float32).reshape(-1).view(np.uint8)
    return buf


def unpack_frame(buf, width=FRAME_W, height=FRAME_H):
    n = width * height
    depth = buf[:2 * n].view(np.uint16).reshape(height, width)
    color = buf[2 * n:5 * n].reshape(height, width, 3)
    mask = buf[5 * n:6 * n].reshape(height, width)
    pose = buf[6 * n:6 * n + 64].view(np.float32).reshape(4, 4)
    return depth, color, mask, pose


def frame_offsets(width=FRAME_W, height=FRAME_H):
    n = width * height
    return 0, 2 * n, 5 * n, 6 * n


def composite_keys(keys, group=None):
    """In-place MIN all-reduce of per-ray keys (int64 view of the u64 keys; all valid keys are < 2^63
    because t > 0, and the 'no hit' key is mapped to INT64_MAX before the reduction)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


NO_HIT = np.int64(np.iinfo(np.int64).max)


def keys_to_int64(keys_u64):
    """u64 device keys -> int64 with UINT64_MAX (no hit) mapped to INT64_MAX so MIN works on signed ints."""
    import torch
    k = keys_u64.view(torch.int64)
    return torch.where(k < 0, torch.full_like(k, int(NO_HIT)), k)


def shard_halo(voxel):
    """Planes a slab must store beyond its owned range for the exact sharded ray-cast:
    ceil(voxel.x / voxel.z) + 2 (previous sample one step back + trilinear tap + refinement)."""
    return int(np.ceil(float(voxel[0]) / float(voxel[2]))) + 2


def stored_range(z0, nz, dz, halo, align=1):
    """Owned planes [z0, z0+nz) -> stored planes including the halo, clipped to the volume.  `align` > 1 rounds the
    stored range outwards to multiples of `align` planes (a stored plane count that is a multiple of 4 keeps the
    integrate kernels on their 128-bit path; a multiple of 8 also keeps whole bricks)."""
    lo = max(0, z0 - halo)
    hi = min(dz, z0 + nz + halo)
    if align > 1:
        lo = (lo // align) * align
        hi = min(dz, -(-hi // align) * align)
    return lo, hi - lo


class SlabVolume:
    """This rank's slab of a volume sharded over `world` ranks.  With `halo` > 0 the handle stores
    (and redundantly integrates) that many extra planes on both sides of the owned range, which is
    what the exact sharded ray-cast needs; no halo data is ever exchanged."""

    def __init__(self, dims, bins, rank, world, device=0, width=FRAME_W, height=FRAME_H, halo=0, plan=None, **kw):
        from .tsdf import Volume
        self.rank, self.world = rank, world
        self.z0, self.nz = plan[rank] if plan is not None else slab_range(rank, world, dims[2])
        sz0, snz = stored_range(self.z0, self.nz, dims[2], halo)
        self.vol = Volume(dims=dims, bins=bins, width=width, height=height, device=device, slab=(sz0, snz),
                          own=(self.z0, self.nz), **kw)
        self.width, self.height = width, height
        try:  # collectives and kernels must share a stream: run the handle on torch's current stream
            import torch
            if torch.cuda.is_available():
                self.vol.set_stream(torch.cuda.current_stream(torch.device("cuda", device)).cuda_stream)
        except ImportError:
            pass

    @classmethod
    def wrap(cls, vol, rank, world, own, width=FRAME_W, height=FRAME_H):
        """A SlabVolume around an existing slab handle (`own` = (z0, nz) of the planes it owns)."""
        self = cls.__new__(cls)
        self.rank, self.world = rank, world
        self.z0, self.nz = own
        self.vol = vol
        self.width, self.height = width, height
        self.replica = None
        return self

    def raycast_sharded(self, s2w, c, w, h, group=None):
        """Exact sharded ray-cast: returns the composited int64 key image (CUDA tensor, identical on every rank)."""
        import torch
        dev = torch.device("cuda", self.vol.desc.device)
        ev1 = torch.empty(w * h, dtype=torch.int64, device=dev)
        ev2 = torch.empty(w * h, dtype=torch.int64, device=dev)
        keys = torch.empty(w * h, dtype=torch.int64, device=dev)
        self.vol.shard_raycast_stage(1, s2w, c, w, h, None, None, ev1.data_ptr())
        composite_keys(ev1, group)
        self.vol.shard_raycast_stage(2, s2w, c, w, h, ev1.data_ptr(), None, ev2.data_ptr())
        composite_keys(ev2, group)
        self.vol.shard_raycast_stage(3, s2w, c, w, h, ev1.data_ptr(), ev2.data_ptr(), keys.data_ptr())
        composite_keys(keys, group)
        return keys

    # ---- ray-cast after the fusion: SDF replicated, histogram sharded (config 4) -----------------------
    def build_sdf_replica(self, plan, bounds, K=None, Kinv=None, group=None):
        """All-gathers the owned SDF planes of every rank's slab into a full-volume, label-free handle on THIS rank
        (4 bytes per voxel: 4.3 GB at 1024^3) and rebuilds its skip map.  `plan`: [(z0, nz)] of all ranks; `bounds`:
        (start, end, voxel, miu) of the volume.  Call once after the last frame (the reference's viewer runs after
        its frame loop, kernel.cpp:101-107); call again if more frames are fused."""
        import torch
        import torch.distributed as dist
        from .tsdf import Volume
        v = self.vol
        dev = torch.device("cuda", v.desc.device)
        dims = tuple(v.dims)
        if getattr(self, "replica", None) is None:
            self.replica = Volume(dims=dims, bins=0, width=self.width, height=self.height, device=v.desc.device, K=K, Kinv=Kinv)
            self.replica.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            self.replica.set_bounds(*bounds)
        self.plan = list(plan)
        cols = dims[0] * dims[1]
        nmax = max(n for _, n in plan)
        mine = torch.empty(cols * nmax, dtype=torch.float32, device=dev)
        v.sdf_planes_dev(self.z0, self.nz, mine.data_ptr(), True)
        if self.world > 1 and dist.is_initialized():
            everyone = torch.empty(self.world * cols * nmax, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(everyone, mine, group=group)
        else:
            everyone = mine
        for r, (z0, n) in enumerate(plan):
            self.replica.sdf_planes_dev(z0, n, everyone.data_ptr() + r * cols * nmax * 4, False)
        self.replica.rebuild_skip_map()
        return self.replica

    def raycast_replicated(self, s2w, c, w, h, group=None):
        """One view on the replicated SDF: this rank marches its share of the image -- every world-th 4-row tile row, so
        that all ranks see the same mix of cheap and expensive regions (the single-volume march: same kernel, same SDF
        bits) --, the hit positions are all-gathered (16 bytes per ray), every rank labels the hits that fall into the
        planes it owns, and one MIN all-reduce composites the keys.  Two collectives per view, no ray is marched twice.
        Returns the composited int64 key image (identical on every rank)."""
        import torch
        import torch.distributed as dist
        dev = getattr(self, "_host_device", None) or torch.device("cuda", self.vol.desc.device)  # (CPU stand-ins in the gloo test)
        prow = self.replica.part_rows(h, self.world)
        key = (w, h)
        if getattr(self, "_ray_buf_key", None) != key:
            self._ray_hits = torch.empty(self.world * prow * w * 4, dtype=torch.float32, device=dev)
            self._ray_part = torch.empty(prow * w * 4, dtype=torch.float32, device=dev)
            self._ray_buf_key = key
        hits, part = self._ray_hits, self._ray_part
        multi = self.world > 1 and dist.is_initialized()
        mine = part if multi else hits
        self.replica.raycast_part_dev(s2w, c, w, h, self.rank, self.world, mine.data_ptr())
        if multi:
            dist.all_gather_into_tensor(hits, part, group=group)
        keys = torch.empty(w * h, dtype=torch.int64, device=dev)
        self.vol.label_hits_parts_dev(hits.data_ptr(), w, h, self.world, keys.data_ptr())
        composite_keys(keys, group)
        return keys

    def set_bounds(self, *a, **k):
        self.vol.set_bounds(*a, **k)

    def broadcast_frame(self, packed_dev, src=0, group=None):
        """`packed_dev`: CUDA uint8 tensor of frame_nbytes(); valid on `src`, overwritten elsewhere
        (ncclBroadcast on the current stream).  Poses are not taken from the buffer: every rank reads
        the (tiny) groundtruth.txt itself, so no device->host read-back is needed per frame."""
        import torch.distributed as dist
        if self.world > 1:
            dist.broadcast(packed_dev, src=src, group=group)
        return packed_dev

    def integrate_packed(self, packed_dev, extrinsic2init):
        """Integrate a packed frame that is resident in this rank's HBM into this rank's slab."""
        o_d, o_c, o_m, _ = frame_offsets(self.width, self.height)
        p = packed_dev.data_ptr()
        self.vol.integrate_dev(p + o_d, p + o_c, p + o_m, extrinsic2init)

    def fuse_packed_sharded(self, packed_dev, extrinsic2init, group=None):
        """Labelled fusion of one frame into a sharded volume = TSDF::launch_kernel (tsdf.cu:418-504) over
        z-slabs: exact sharded march from the incoming camera (three MIN all-reduces), fold of the hits
        this rank owns into the overlap tables, SUM all-reduce of the integer tables (exact, so every
        rank decides on the single-GPU tables), decision + relabel on every rank, integrate.  The label
        image inside `packed_dev` is relabelled in place; returns (lut, report) -- (None, None) for the
        first frame, which only fixes num_objs (tsdf.cu:464-467)."""
        import torch
        import torch.distributed as dist
        v = self.vol
        o_d, o_c, o_m, _ = frame_offsets(self.width, self.height)
        p = packed_dev.data_ptr()
        lut = rep = None
        if v.bins > 0 and v.info().n_obs > 0:
            dev = packed_dev.device
            n = self.width * self.height
            ev1 = torch.empty(n, dtype=torch.int64, device=dev)
            ev2 = torch.empty(n, dtype=torch.int64, device=dev)
            keys = torch.empty(n, dtype=torch.int64, device=dev)
            v.shard_backproj_stage(1, extrinsic2init, None, None, ev1.data_ptr())
            composite_keys(ev1, group)
            v.shard_backproj_stage(2, extrinsic2init, ev1.data_ptr(), None, ev2.data_ptr())
            composite_keys(ev2, group)
            v.shard_backproj_stage(3, extrinsic2init, ev1.data_ptr(), ev2.data_ptr(), keys.data_ptr())
            composite_keys(keys, group)
            n64, ntot = v.fold_table_bytes()
            tables = torch.empty(ntot, dtype=torch.uint8, device=dev)
            v.shard_fold(p + o_m, keys.data_ptr(), self.rank == 0, tables.data_ptr())
            if self.world > 1 and dist.is_initialized():
                dist.all_reduce(tables[:n64].view(torch.int64), op=dist.ReduceOp.SUM, group=group)
                dist.all_reduce(tables[n64:].view(torch.int32), op=dist.ReduceOp.SUM, group=group)
            lut, rep = v.shard_merge_finish(tables.data_ptr(), p + o_m)
        elif v.bins > 0:
            v.shard_first_frame(p + o_m)
        v.integrate_dev(p + o_d, p + o_c, p + o_m, extrinsic2init)
        return lut, rep

    def close(self):
        if getattr(self, "replica", None) is not None:
            self.replica.close()
            self.replica = None
        self.vol.close()
