"""torchrun --nproc-per-node N tools/dp_breakdown.py : where does a data-parallel step spend its time?"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import bench  # noqa: E402
from photo_slam_b200 import _lib, trainer as T  # noqa: E402


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    class A: points = 3_000_000; camera = "replica"
    world, rank, local = bench.dist_setup(A)
    dev = torch.device("cuda", local)
    scene, cam, host, devcam, gt = bench.make_inputs(A, rank, dev)
    model = T.GaussianModel.from_numpy(scene, dev)
    model.trainingSetup(T.GaussianOptimizationParams())
    tr = T.DataParallelTrainer(model)
    tr.trainForOneIteration(devcam, gt); tr.result()
    m = model
    cs = tr._cstep(True)
    cm, cc = m._cmodel(), T._ccamera(devcam)
    ptrs = (C.c_void_p * 6)(*[s.data_ptr() for s in tr.segs])
    st = lambda: torch.cuda.current_stream().cuda_stream

    def bwd():
        _lib.check(tr.L.psb_trainer_backward(tr.h, m.num_points(), 16, C.byref(cm), C.byref(cc), tr.background.data_ptr(), gt.data_ptr(), None,
                                             C.byref(cs), None, None, ptrs, st()), "bwd")
    t_b = timed(bwd)
    t_ar = timed(lambda: dist.all_reduce(tr.flat))
    t_ar6 = timed(lambda: [dist.all_reduce(tr.segs[i][o:o + n]) for (i, o, n) in tr._chunks] if hasattr(tr, "_chunks") else None)
    t_ad = timed(lambda: _lib.check(tr.L.psb_adam_update(m.num_points(), 16, C.byref(cm), ptrs, C.byref(cs), 0.5, st()), "adam"))
    t_full = timed(lambda: tr.trainForOneIteration(devcam, gt))
    if rank == 0:
        print(f"world {world}: backward(no update) {t_b:.3f} ms | all_reduce 708MB one call {t_ar:.3f} ms, chunked {t_ar6:.3f} ms | adam_update {t_ad:.3f} ms | full DP step {t_full:.3f} ms", flush=True)
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
