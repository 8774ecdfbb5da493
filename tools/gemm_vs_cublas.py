"""bf16 GEMM at the step's shapes: this repo's kernel vs cuBLAS (torch.matmul) on the same box, burst and sustained, with
NVML clocks / power. usage: python tools/gemm_vs_cublas.py [ncu]   ("ncu": one launch of each at the dominant shape)"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

from llamax_b200 import ops

pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"


class Clocks:
    def __enter__(self):
        self.s, self.p, self.on = [], [], True
        self.t = threading.Thread(target=self.run, daemon=True)
        self.t.start()
        return self

    def run(self):
        while self.on:
            self.s.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3)
            time.sleep(0.01)

    def __exit__(self, *a):
        self.on = False
        self.t.join()

    def summary(self):
        s = sorted(self.s)
        return f"{s[len(s) // 2]} MHz {max(self.p):.0f} W"


def burst(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def sustained(fn, secs=2.0):
    t = burst(fn, 2)
    reps = max(10, int(secs * 1e3 / t))
    with Clocks() as c:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, c.summary()


def main():
    shapes = [(16384, 4096, 28688), (16384, 14336, 4096), (16384, 4096, 6168), (16384, 4096, 4096), (8192, 8192, 8192)]
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":
        M, N, K = shapes[0]
        a = torch.randn(M, K + 48, device=dev).bfloat16()[:, :K]
        b = torch.randn(N, K + 48, device=dev).bfloat16()[:, :K]
        for _ in range(2):
            ops.bf16_gemm(a, b)
            torch.matmul(a, b.t())
        torch.cuda.synchronize()
        print("ok")
        sys.exit(0)

    for (M, N, K) in shapes:
        Kp = (K + 63) // 64 * 64
        a = torch.randn(M, Kp, device=dev).bfloat16()[:, :K]          # 128-byte pitches, as the fused block allocates them
        b = torch.randn(N, Kp, device=dev).bfloat16()[:, :K]
        fl = 2.0 * M * N * K
        for name, fn in (("llamax_b200", lambda: ops.bf16_gemm(a, b)), ("cuBLAS", lambda: torch.matmul(a, b.t()))):
            tb = burst(fn)
            ts, clk = sustained(fn)
            print(f"[{M},{N},{K}] {name:12s} burst {fl / tb / 1e9:7.0f} TF/s  sustained {fl / ts / 1e9:7.0f} TF/s  [{clk}]", flush=True)
        del a, b


if __name__ == "__main__":
    main()
