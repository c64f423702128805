"""Host-side breakdown of the end-to-end loop (bench.py's e2e leg): where the wall time of one iteration goes.
Usage: python tools/e2e_breakdown.py [steps]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    args = argparse.Namespace(points=3_000_000, seed=0, camera="replica")
    dev = torch.device("cuda:0")
    from photo_slam_b200 import trainer as T
    scene, cam, host, devcam, gt_dev = bench.make_inputs(args, 0, dev)
    model = T.GaussianModel.from_numpy(scene, dev)
    model.trainingSetup(T.GaussianOptimizationParams())
    tr = T.GaussianTrainer(model)
    for _ in range(3):
        tr.trainForOneIteration(devcam, gt_dev)
        tr.result()
    hostcam = dict(devcam, viewmatrix=host["viewmatrix"], projmatrix=host["projmatrix"], campos=host["campos"])

    def loop(name, fn, sync_each=False):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(f"{name}: {e0.elapsed_time(e1) / steps:.3f} ms/iter (events), {(t1 - t0) * 1e3 / steps:.3f} ms/iter (wall)", flush=True)

    loop("device inputs, no read-back", lambda: tr.trainForOneIteration(devcam, gt_dev))

    def dev_result():
        tr.trainForOneIteration(devcam, gt_dev)
        tr.result()
    loop("device inputs, loss read every iteration", dev_result)
    loop("trainHost (host inputs + loss read)", lambda: tr.trainHost(hostcam, host["gt"]))

    # host time of the pieces
    enq, res = [], []
    torch.cuda.synchronize()
    for _ in range(steps):
        a = time.perf_counter()
        tr.trainForOneIteration(devcam, gt_dev)
        b = time.perf_counter()
        tr.result()
        c = time.perf_counter()
        enq.append(b - a)
        res.append(c - b)
    torch.cuda.synchronize()
    print(f"host: enqueue of one step median {np.median(enq) * 1e3:.3f} ms (max {np.max(enq) * 1e3:.3f}); result() wait median {np.median(res) * 1e3:.3f} ms", flush=True)
    enq = []
    for _ in range(steps):
        a = time.perf_counter()
        tr.trainForOneIteration(devcam, gt_dev)
        enq.append(time.perf_counter() - a)
        torch.cuda.synchronize()
    print(f"host: enqueue on an idle GPU median {np.median(enq) * 1e3:.3f} ms", flush=True)




def probe():
    """Why is a loop that reads the loss every iteration slower than the free-running loop?"""
    import pynvml
    pynvml.nvmlInit()
    nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    clock = lambda: pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
    args = argparse.Namespace(points=3_000_000, seed=0, camera="replica")
    dev = torch.device("cuda:0")
    from photo_slam_b200 import trainer as T
    scene, cam, host, devcam, gt_dev = bench.make_inputs(args, 0, dev)
    model = T.GaussianModel.from_numpy(scene, dev)
    model.trainingSetup(T.GaussianOptimizationParams())
    tr = T.GaussianTrainer(model)
    for _ in range(5):
        tr.trainForOneIteration(devcam, gt_dev)
        tr.result()
    torch.cuda.synchronize()
    s = torch.cuda.current_stream()

    snap, it0 = model.snapshot(), tr.iteration

    def run(name, after, n=150, profile=False):
        model.restore(snap)                      # every mode times the same 150 iterations
        tr.iteration = it0
        tr.set_profiling(profile)
        torch.cuda.synchronize()
        clk = []
        t0 = time.perf_counter()
        for i in range(n):
            tr.trainForOneIteration(devcam, gt_dev)
            after()
            if i % 25 == 24:
                clk.append(clock())
        s.synchronize()
        dt = (time.perf_counter() - t0) * 1e3 / n
        extra = ""
        if profile:
            st = tr.stage_times()
            extra = "; last step: " + ", ".join(f"{k} {v:.3f}" for k, v in st.items()) + f"; sum {sum(st.values()):.3f}"
        print(f"{name}: {dt:.3f} ms/iter, SM clock {clk}{extra}", flush=True)

    run("free-running", lambda: None)
    run("free-running, profiled", lambda: None, profile=True)
    run("stream.synchronize() each iteration", s.synchronize)
    run("stream.synchronize() each iteration, profiled", s.synchronize, profile=True)
    run("result() each iteration", tr.result)
    run("result() each iteration, profiled", tr.result, profile=True)

    def sleepy():
        time.sleep(0.0028)
        tr.result()
    run("sleep 2.8 ms then result()", sleepy)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "probe":
        probe()
    else:
        main()
