"""Streaming ceiling check: the flat Adam kernel (4 read + 3 write streams, 128-bit accesses) and a device copy on
tensors of the f_rest size, to tell what HBM delivers for the fused backward+Adam kernel's access mix."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from photo_slam_b200 import trainer as T  # noqa: E402


def main():
    L = T._bind()
    n = 3_000_000 * 45
    dev = torch.device("cuda:0")
    p, m, v, g = (torch.rand(n, device=dev) for _ in range(4))
    s = T._Step()
    for i in range(6):
        s.lr[i] = 1e-3
    s.beta1, s.beta2, s.eps, s.step, s.lambda_dssim, s.sh_degree, s.update_densify_stats = 0.9, 0.999, 1e-15, 5, 0.2, 3, 0
    L.psb_adam_flat.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
    st = torch.cuda.current_stream().cuda_stream

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms = t(lambda: L.psb_adam_flat(n, p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), 1e-3, C.byref(s), 1.0, st))
    print(f"adam_flat  n={n}: {ms:.3f} ms, {7 * 4 * n / ms / 1e6:.0f} GB/s (4 read + 3 write streams)")
    q = torch.empty_like(p)
    ms = t(lambda: q.copy_(p))
    print(f"copy       n={n}: {ms:.3f} ms, {2 * 4 * n / ms / 1e6:.0f} GB/s")
    ms = t(lambda: torch.add(p, m, out=q))
    print(f"add        n={n}: {ms:.3f} ms, {3 * 4 * n / ms / 1e6:.0f} GB/s (2 read + 1 write)")
    ms = t(lambda: p.mul_(1.0001))
    print(f"mul_       n={n}: {ms:.3f} ms, {2 * 4 * n / ms / 1e6:.0f} GB/s (in place)")
    ms = t(lambda: torch._foreach_mul_([p, m, v], 1.0001))
    print(f"foreach mul_ x3 (3 in-place streams): {ms:.3f} ms, {6 * 4 * n / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
