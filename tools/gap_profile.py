"""How much of the step is NOT inside a kernel? Two measurements on the bench's text workload (8B shape, 8 x 2048):
(1) host time to enqueue one step (no synchronisation inside) against the step's device time: if the host finishes far
    ahead, the launch queue is never empty and only the hardware's kernel-to-kernel turnaround is left;
(2) a CUPTI kernel trace of one step (torch.profiler): sum of the idle intervals between consecutive kernels, and the
    kernels after which the longest ones occur.
(3) device time per kernel name from the same trace (library and torch kernels included).
Usage: python tools/gap_profile.py [layers] [text|speech]      (run under gpurun; prints tables)
"""
import os
import sys
import time
from collections import defaultdict
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    workload = sys.argv[2] if len(sys.argv) > 2 else "text"
    args = SimpleNamespace(layers=layers, weight_only=False, rank=8, batch=8, seq=2048)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    model, cfg = bench.build_model(args, dev, workload)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.0, fused=True)
    host, positions, _ = bench.make_batch(args, cfg, 0, workload)
    batch = {k: v.to(dev) for k, v in host.items()}

    def step():
        if workload == "text":
            loss = model(batch["tokens"], labels=batch["labels"], block_mask=None)
        else:
            loss = model(batch["audio"], batch["tokens"], labels=batch["labels"], prefix_lm=True)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # (1) host enqueue time vs device time
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        step()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"host enqueue {1e3 * (t1 - t0):8.1f} ms   device {e0.elapsed_time(e1):8.1f} ms")
    # (2) kernel trace of one step
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    evs.sort(key=lambda e: e.time_range.start)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    span = evs[-1].time_range.end - evs[0].time_range.start
    gaps = defaultdict(lambda: [0, 0.0, 0.0])
    hist = defaultdict(int)
    total_gap = 0.0
    end = evs[0].time_range.end
    prev = evs[0]
    for e in evs[1:]:
        g = e.time_range.start - end
        if g > 0:
            total_gap += g
            key = (prev.name[:48], e.name[:48])
            s = gaps[key]
            s[0] += 1
            s[1] += g
            s[2] = max(s[2], g)
            hist[min(int(g), 20)] += 1
        if e.time_range.end > end:
            end, prev = e.time_range.end, e
    print(f"kernels {len(evs)}   span {span / 1e3:.2f} ms   busy (sum of durations) {busy / 1e3:.2f} ms   "
          f"idle between kernels {total_gap / 1e3:.2f} ms  ({100 * total_gap / span:.2f} %)")
    print("gap histogram (us -> count):", dict(sorted(hist.items())))
    by_name = defaultdict(lambda: [0, 0.0])
    for e in evs:
        d = by_name[e.name[:90]]
        d[0] += 1
        d[1] += e.time_range.end - e.time_range.start
    print("device time per kernel name (top 40):")
    for name, (n, tot) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {tot / 1e3:8.3f} ms  {100 * tot / busy:5.2f} %  n={n:5d}  {name}")
    print("largest idle totals by (previous kernel -> next kernel):")
    for (a, b), (n, tot, mx) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  {tot / 1e3:7.3f} ms  n={n:5d}  max {mx:7.1f} us   {a}  ->  {b}")


if __name__ == "__main__":
    main()
