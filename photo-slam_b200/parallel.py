"""Host-side logic of the keyframe-sharded data-parallel path. Device-agnostic (torch tensors on
any device, any torch.distributed backend): exercised with gloo on CPU in tests/test_parallel_cpu.py and with NCCL
on the B200s by trainer.DataParallelTrainer / bench.py --gpus N."""
import torch

SIZES = (3, 3, 45, 1, 3, 4)   # floats per Gaussian of xyz, features_dc, features_rest, opacity, scaling, rotation
PER_GAUSSIAN = sum(SIZES)     # 59


def shard_schedule(schedule, rank, world):
    """Rank r consumes entries r, r+K, r+2K, ... of the shuffled keyframe schedule: K consecutive entries of the
    reference's single-view schedule (gaussian_mapper.cpp:1126-1173) form one data-parallel step."""
    usable = len(schedule) - len(schedule) % world
    return list(schedule[rank:usable:world])


CHUNK = 128                   # Gaussians per ownership chunk of the NVLink data-parallel step (= one backward block)


def owner_of_rows(P, world, device="cpu"):
    """int64 [P]: rank that owns each Gaussian's optimizer state in the p2p step (psb_dp_*): chunk c = row // 128 belongs
    to rank c % world — interleaved, so the records every rank pushes spread evenly over all NVLink destinations at any
    moment of the backward kernel (contiguous shards would make all ranks hit the same owner at the same time)."""
    return (torch.arange(P, device=device) // CHUNK) % world


def gather_owned_rows(tensors, rank, world, group=None):
    """In place: every tensor [P, ...] ends up with row r taken from the rank that owns r. Implemented as zero the rows of
    other owners + all-reduce(SUM) — rare (before densification / a checkpoint), any backend."""
    import torch.distributed as dist
    if world == 1:
        return
    for t in tensors:
        mine = (owner_of_rows(t.size(0), world, t.device) == rank).view(-1, *([1] * (t.dim() - 1)))
        t.mul_(mine.to(t.dtype))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def _pad4(n):
    return -(-n // 4) * 4


class GradBuffer:
    """ONE flat float32 buffer with six views in the reference's tensor shapes, so the whole gradient of a step is a single
    all-reduce call. Every segment starts on a 16-byte boundary (the kernels store the features_rest and rotation gradients
    with 128-bit accesses; P need not be a multiple of 4): [P*59 (+ up to 3 pad floats per segment)]."""

    def __init__(self, P, device):
        self.P = P
        self.flat = torch.zeros(sum(_pad4(P * s) for s in SIZES), dtype=torch.float32, device=device)
        self.segments, o = [], 0
        for s in SIZES:
            self.segments.append(self.flat[o:o + P * s])
            o += _pad4(P * s)

    def views(self):
        P = self.P
        shapes = [(P, 3), (P, 1, 3), (P, 15, 3), (P, 1), (P, 3), (P, 4)]
        return [seg.view(*sh) for seg, sh in zip(self.segments, shapes)]

    def all_reduce(self, group=None):
        """Sum over ranks; returns the factor the optimizer applies (mean over the K views of the step)."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world


def reduce_densify_stats(max_radii2D, xyz_gradient_accum, denom, group=None):
    """Called once, right before densify/prune: max over ranks for the radii, sum for the accumulators
    (every rank then holds the statistics of all views seen since the last densification)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(denom, op=dist.ReduceOp.SUM, group=group)


class SlabGradBuffer:
    """The same [P*59] floats, laid out slab-major: Gaussians are cut into `nslabs` contiguous ranges (boundaries at
    multiples of 128) and each slab keeps its six gradient blocks back to back, so ONE all-reduce call per slab covers
    all parameters of those Gaussians and can start as soon as the slab's backward kernel has finished."""

    def __init__(self, P, device, nslabs=4):
        self.P = P
        per = -(-P // nslabs)
        per = -(-per // 128) * 128
        self.slabs = []           # (first, count)
        f = 0
        while f < P:
            self.slabs.append((f, min(per, P - f)))
            f += per
        self.flat = torch.zeros(max(sum(self._slab_floats(c) for _, c in self.slabs), 4), dtype=torch.float32, device=device)

    def _slab_floats(self, count):
        return sum(_pad4(count * k) for k in SIZES)

    def region(self, s):
        o = sum(self._slab_floats(c) for _, c in self.slabs[:s])
        return self.flat[o:o + self._slab_floats(self.slabs[s][1])]

    def blocks(self, s):
        """Six 1-D views (one per parameter tensor) of slab s, each starting on a 16-byte boundary."""
        first, count = self.slabs[s]
        out, o = [], sum(self._slab_floats(c) for _, c in self.slabs[:s])
        for k in SIZES:
            out.append(self.flat[o:o + count * k])
            o += _pad4(count * k)
        return out

    def kernel_pointers(self, s):
        """Base addresses such that row g of tensor i lands at blocks(s)[i][(g - first) * SIZES[i]] (see psb_trainer_backward_slab)."""
        first, _ = self.slabs[s]
        return [b.data_ptr() - first * k * 4 for b, k in zip(self.blocks(s), SIZES)]
