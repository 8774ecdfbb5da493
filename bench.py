#!/usr/bin/env python
"""bench.py — train tokens/s of Llama-3.1-8B (frozen INT8 base + LoRA r=8) on B200, BASELINE.json `configs[1]`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload text|speech]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one full optimizer step over one synthetic batch: forward, backward
through all 32 decoder blocks + LM head/cross-entropy, gradient all-reduce (N > 1), fused AdamW on the trainable
parameters.  `value` is whole-job positions/s with the batch resident in HBM; `e2e` repeats the measurement through
the model's public forward with the batch in pinned host memory (H2D inside the timed region) and a D2H read of the
loss every step.  `--impl reference` times the CPU oracle port of the reference path on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train tokens/s, Llama-3.1-8B INT8+LoRA prefix-LM"
UNIT = "tokens/s"

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# (profiles/r1_ncu_full_hot_kernels.csv, kernel ID 0; earlier captures of the same kernel: 5.2-7.1 GB), keyed by GEMM shape
NCU_TRAFFIC_BYTES = {"[M=16384,N=4096,K=28688]": 5247264000 + 131630000}  # algorithmic: 1.31 GB (A re-read per n-sweep)

LLAMA8B = dict(embed_dim=4096, num_layers=32, head_dim=128, num_heads=32, num_kv_heads=8, intermediate_dim=14336,
               vocab_size=128256, rope_base=500000, is_llama3_1=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="text", choices=["text", "speech"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--weight-only", action="store_true", help="weight-only INT8 (bf16 GEMM) instead of dynamic")
    ap.add_argument("--rank", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cta-group", type=int, default=2)
    ap.add_argument("--int8-grad-input", action="store_true",
                    help="OPT-IN, NON-PARITY: grad_input on the int8 tensor path (SURVEY 8 f4); never the headline")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        # median over the upper half of samples = under load (idle samples at the edges are dropped)
        busy = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ ours
def build_model(args, device):
    import torch
    from llamax_b200.modelling import AudioConfig, Llama, LlamaAudio, LlamaConfig, apply_linear_adapter_
    from llamax_b200.subclasses import quantize_linear_

    cfgd = dict(LLAMA8B)
    cfgd["num_layers"] = args.layers
    cfgd["max_seq_len"] = 4096
    cfg = LlamaConfig(**cfgd)
    torch.manual_seed(1234)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = LlamaAudio(cfg, AudioConfig(n_mels=80)) if args.workload == "speech" else Llama(cfg)
    finally:
        torch.set_default_dtype(prev)
    quantize_linear_(model.layers, "int8", dynamic_int8_act=not args.weight_only)
    apply_linear_adapter_(model.layers, "lora", rank=args.rank)
    g = torch.Generator(device=device).manual_seed(7)
    for m in model.modules():
        if hasattr(m, "lora_b"):  # non-zero B so that every backward GEMM does real work
            m.lora_b.data.normal_(0, 0.02, generator=g)
    model.tok_embeddings.requires_grad_(False)   # small-payload DP: only LoRA, norms (+ conv stem) train
    model.output.requires_grad_(False)
    model.build_cache()
    return model, cfg


def make_batch(args, cfg, rank, device):
    """Synthetic MetaMathQA-shaped (text) or LibriSpeech-shaped (speech) batch in pinned host memory."""
    import torch

    g = torch.Generator().manual_seed(100 + rank)
    if args.workload == "text":
        tokens = torch.randint(0, cfg.vocab_size, (args.batch, args.seq), generator=g)
        labels = torch.roll(tokens, -1, 1)
        labels[:, : args.seq // 4] = -100   # prompt positions carry no loss
        labels[:, -1] = -100
        host = dict(tokens=tokens.pin_memory(), labels=labels.pin_memory())
        positions = args.batch * args.seq
    else:
        T = 256
        audio = torch.randn(args.batch, 480000, generator=g)  # 30 s @ 16 kHz -> 1500 prefix positions
        tokens = torch.randint(0, cfg.vocab_size, (args.batch, T), generator=g)
        labels = torch.roll(tokens, -1, 1)
        labels[:, -1] = -100
        host = dict(audio=audio.pin_memory(), tokens=tokens.pin_memory(), labels=labels.pin_memory())
        positions = args.batch * (1500 + T)
    n_label = int((host["labels"] != -100).sum())
    return host, positions, n_label


def run_ours(args):
    import torch
    import torch.distributed as dist

    from llamax_b200 import ops
    from llamax_b200.dp import GradBucket, init_distributed
    from llamax_b200.modelling import PrefixLM

    rank, world, local_rank = init_distributed()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    ops.set_gemm_cta_group(args.cta_group)
    if args.int8_grad_input:
        from llamax_b200.modelling import fused_block as _fb

        _fb.set_int8_grad_input(True)
    model, cfg = build_model(args, device)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.0, fused=True)
    bucket = GradBucket(params)
    host, positions, n_label = make_batch(args, cfg, rank, device)
    resident = {k: v.to(device) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def step(batch):
        if args.workload == "text":
            loss = model(batch["tokens"], labels=batch["labels"], block_mask=None)
        else:
            loss = model(batch["audio"], batch["tokens"], labels=batch["labels"], prefix_lm=True)
        loss.backward()
        bucket.allreduce_()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n):
            if e2e:
                batch = {k: v.to(device, non_blocking=True) for k, v in host.items()}
                last = step(batch).item()          # D2H read of the loss, every step
            else:
                last = step(resident)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, (last if e2e else last.item())

    for _ in range(max(args.warmup, 3)):
        step(resident)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # timed region: only the dominant kernel class (bf16 GEMM, a few hundred launches) carries CUDA-event pairs, so the
    # instrumentation costs nothing measurable; the full per-kernel breakdown comes from a separate instrumented pass
    ops.TIMING.enable(only={"bf16_gemm"})
    ms, loss_val = timed(args.steps, e2e=False)
    kern_dom = ops.TIMING.summary()        # bf16 GEMM launches of the timed region (roofline object)
    ops.TIMING.disable()
    ms_e2e, loss_e2e = timed(args.steps, e2e=True)
    ops.TIMING.enable()
    ms_prof, _ = timed(args.steps, e2e=False)
    kern = ops.TIMING.summary()            # every library launch, by kernel class (events around each: ~2 % overhead)
    ops.TIMING.disable()
    kern.update({k: v for k, v in kern_dom.items()})   # the bf16 GEMM entries quoted are those of the timed region
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ms_step = ms / args.steps
    value = positions * world / (ms_step / 1e3)
    e2e_value = positions * world / (ms_e2e / args.steps / 1e3)
    # dominant kernel = the bf16 tcgen05 GEMM; the roofline object is quoted on its heaviest shape (the w1|w3
    # grad_input GEMM, contraction over 2F + LoRA columns), timed live by CUDA events around each launch
    shaped = {k: v for k, v in kern.items() if k.startswith("bf16_gemm[")}
    kern = {k: v for k, v in kern.items() if "[" not in k}
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    roof = None
    if shaped:
        name, dom = max(shaped.items(), key=lambda kv: kv[1]["ms"])
        ach = dom["flops"] / (dom["ms"] / 1e3) / 1e12
        what = "grad_input of w1|w3" if "N=4096,K=28" in name else "largest bf16 GEMM of the step"
        roof = {"kernel": "gemm_kernel<bf16,cta_group::%d,rank0> %s (%s)" % (args.cta_group, name[len("bf16_gemm"):], what),
                "bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1400",
                "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4),
                "traffic": NCU_TRAFFIC_BYTES.get(name[len("bf16_gemm"):]),
                "flops_per_launch": dom["flops"] / dom["n"], "us_per_launch": round(1e3 * dom["ms"] / dom["n"], 1),
                "launches_per_step": dom["n"] // args.steps, "ms_per_step": round(dom["ms"] / args.steps, 2),
                "all_bf16_gemm_tflops": round(kern["bf16_gemm"]["flops"] / (kern["bf16_gemm"]["ms"] / 1e3) / 1e12, 1)}
    shares = {k: {"ms_per_step": round(v["ms"] / args.steps, 2), "launches_per_step": v["n"] // args.steps,
                  **({"achieved_tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)} if v.get("flops") else {}),
                  **({"achieved_gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1)} if v.get("bytes") else {})}
              for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8 fwd (int32 acc) / bf16 bwd" if not args.weight_only else "bf16 (int8 weights)",
        "data": "synthetic",
        "config": {"workload": workload_label(args, positions),
                   "layers": args.layers, "global_batch": args.batch * world,
                   "seq_len": args.seq if args.workload == "text" else positions // args.batch,
                   "int8_mode": "weight-only" if args.weight_only else "dynamic_int8_act", "parallelism": f"dp{world}",
                   **({"NON_PARITY_OPT_IN": "int8 grad_input (gradients quantised to 8 bit per row): not the reference's "
                                            "numerics, not a headline number"} if args.int8_grad_input else {}),
                   "l2": "per-step working set (>10 GB activations + 7 GB weights) far exceeds the 126 MB L2; no flush needed",
                   "positions_per_step": positions * world, "label_tokens_per_step": n_label * world},
        "label_tokens_per_s": round(n_label * world / (ms_step / 1e3), 1),
        "loss": round(float(loss_val), 4),
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 2)},
        "gpu_launches": sum(v["n"] * (4 if k == "attn_bwd" else 2 if k == "lora_wgrad" else 1) for k, v in kern.items()),
        "roofline": roof,
        "kernels": shares,
        "kernels_note": "per-class device time from a separate fully instrumented pass of the same steps (%.2f ms/step); "
                        "bf16_gemm and the roofline object are from the timed region itself" % (ms_prof / args.steps),
        "dp_payload_bytes": bucket.nbytes(),
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_reference(args, steps=1, warmup=1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def workload_label(args, positions=None):
    """`config.workload` of both arms (the reference arm times a bounded sample of the same workload)."""
    seq = args.seq if args.workload == "text" else (positions // args.batch if positions else 1500 + 256)
    return ("Llama-3.1-8B shape, frozen INT8 base + LoRA r=%d, %s, seq %d batch %d per GPU" %
            (args.rank, "MetaMathQA-shaped text SFT (causal)" if args.workload == "text"
             else "LibriSpeech-shaped 30 s audio prefix (1500) + 256 text, prefix-LM", seq, args.batch))


# ------------------------------------------------------------------------------------------------ reference (CPU oracle port)
def cpu_reference(args, steps=1, warmup=0):
    """Time the oracle port of the reference decoder path on the host cores: ONE 8B-shape decoder block fwd+bwd
    (dynamic INT8 + LoRA r=8, reference op sequence in bf16) on `sample_tokens` positions, + the LM head on the same
    positions; tokens/s is extrapolated to the 32-block model: tokens / (32 * t_block + t_head)."""
    import torch

    from oracle import ref_ops as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    S = 512
    D, Fd, Hq, Hkv, hd = 4096, 14336, 32, 8, 128
    lw = R.LayerWeights()
    shapes = dict(wq=(Hq * hd, D), wk=(Hkv * hd, D), wv=(Hkv * hd, D), wo=(D, Hq * hd), w1=(Fd, D), w3=(Fd, D), w2=(D, Fd))
    for n, (o, i) in shapes.items():
        lw.w8[n] = torch.randint(-127, 128, (o, i), dtype=torch.int8)
        lw.ws[n] = (torch.rand(o) * 1e-3).bfloat16()
        lw.lora_a[n] = (torch.randn(args.rank, i) * 0.02).bfloat16().requires_grad_(True)
        lw.lora_b[n] = (torch.randn(o, args.rank) * 0.02).bfloat16().requires_grad_(True)
    lw.attention_norm = torch.ones(D, dtype=torch.bfloat16, requires_grad=True)
    lw.ffn_norm = torch.ones(D, dtype=torch.bfloat16, requires_grad=True)
    rope = R.build_rope(hd, S, 500000, True)
    x = torch.randn(1, S, D).bfloat16().requires_grad_(True)
    w_out = (torch.randn(128256, D) * 0.02).bfloat16()
    labels = torch.randint(0, 128256, (S,))
    dynamic = not args.weight_only

    def one():
        t0 = time.perf_counter()
        out = R.transformer_layer_ref(x, rope, lw, Hq, Hkv, hd, 0, dynamic)
        out.backward(torch.ones_like(out))
        t1 = time.perf_counter()
        h = x.detach().reshape(S, D).requires_grad_(True)
        loss = torch.nn.functional.cross_entropy((h @ w_out.T).float(), labels)
        loss.backward()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    for _ in range(warmup):
        one()
    best = None
    for _ in range(max(steps, 1)):
        tb, th = one()
        tot = 32 * tb + th
        best = tot if best is None or tot < best else best
    return {"value": round(S / best, 2), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (PyTorch CPU, bf16, {'dynamic' if dynamic else 'weight-only'} INT8 + LoRA r={args.rank}): one 8B-shape "
                      f"decoder block fwd+bwd + LM head/CE on 1x{S} positions, extrapolated x32 blocks",
            "seconds_per_sample": round(best / 32, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    base = cpu_reference(args, steps=max(1, min(args.steps, 3)), warmup=1)  # first call pays thread-pool and allocator start-up
    wall = time.perf_counter() - t0
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * base["seconds_per_sample"] * 32, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16/int8 (CPU)",
           "data": "synthetic",
           "config": {"workload": workload_label(args), "layers": args.layers, "global_batch": args.batch,
                      "seq_len": args.seq if args.workload == "text" else 1756,
                      "int8_mode": "weight-only" if args.weight_only else "dynamic_int8_act", "parallelism": "cpu",
                      "sample": "bounded: one decoder block + LM head on 1x512 positions per step, extrapolated to 32 blocks"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": round(wall, 1)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
