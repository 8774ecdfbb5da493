"""GPU parity of the data-parallel path (last row): 2-GPU DP gradients over NCCL == 1-GPU gradients on
the concatenated batch. Needs two CUDA devices (`gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py -m gpu`);
skipped on a single-GPU box (the world-size-2 logic is also covered on CPU by tests/test_dp_gloo.py)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_path: str):
    import torch.distributed as dist

    from llamax_b200.dp import GradBucket
    from llamax_b200.modelling import PrefixLM
    from tests.helpers import build_tiny_llama, rel_err

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        P, B, S = 48, 4, 160
        torch.manual_seed(17)
        tokens = torch.randint(0, 1024, (B, S))
        labels = torch.randint(0, 1024, (B, S))       # every position carries a label: equal counts per rank, so the
        #                                               mean of the per-rank mean losses is the global mean loss

        def grads_of(tok, lab):
            model = build_tiny_llama(True, num_layers=2).cuda()       # same seed on every rank: identical replicas
            model.build_cache()
            model.tok_embeddings.requires_grad_(False)
            model.output.requires_grad_(False)
            params = [p for p in model.parameters() if p.requires_grad]
            loss = model(tok.cuda(), labels=lab.cuda(), block_mask=PrefixLM(P))
            loss.backward()
            return model, params, loss

        per = B // world
        model, params, loss = grads_of(tokens[rank * per : (rank + 1) * per], labels[rank * per : (rank + 1) * per])
        bucket = GradBucket(params)
        bucket.allreduce_()                       # ONE ncclAllReduce(AVG) over the flat bf16 bucket
        torch.cuda.synchronize()
        lsum = loss.detach().float().clone()
        dist.all_reduce(lsum, op=dist.ReduceOp.AVG)
        if rank == 0:
            _, ref_params, ref_loss = grads_of(tokens, labels)     # single GPU, concatenated batch
            assert len(ref_params) == len(params)
            worst = max(rel_err(p.grad, q.grad) for p, q in zip(params, ref_params))
            # both sides are bf16 gradients of the same kernels; they differ by the bf16 rounding of the per-rank
            # gradients before the average and by the order of the fp32 reductions inside the kernels
            assert worst <= 1e-2, worst
            assert abs(lsum.item() - ref_loss.item()) <= 2e-3 * abs(ref_loss.item()), (lsum.item(), ref_loss.item())
            with open(out_path, "w") as f:
                f.write(f"ok worst_grad_rel_err={worst:.3e} loss_dp={lsum.item():.5f} loss_1gpu={ref_loss.item():.5f} "
                        f"params={len(params)} payload_bytes={bucket.nbytes()}\n")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_dp_grads_equal_single_gpu_grads(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp

    out = str(tmp_path / "dp2.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    line = open(out).read()
    print(line)
    assert line.startswith("ok")
