"""CPU, world_size 2 over gloo: the data-parallel gradient bucket averages trainable grads across ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from llamax_b200.dp import GradBucket, init_distributed

    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3, dtype=torch.bfloat16)), torch.nn.Parameter(torch.zeros(7, dtype=torch.bfloat16)),
              torch.nn.Parameter(torch.zeros(2), requires_grad=False)]
    params[0].grad = torch.full((5, 3), float(rank + 1), dtype=torch.bfloat16)
    params[1].grad = None if rank == 0 else torch.full((7,), 4.0, dtype=torch.bfloat16)  # a rank may miss a grad
    bucket = GradBucket(params)
    assert bucket.numel == 22 and bucket.nbytes() == 44
    bucket.allreduce_()
    ok = torch.allclose(params[0].grad.float(), torch.full((5, 3), 1.5)) and torch.allclose(params[1].grad.float(), torch.full((7,), 2.0))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_single_process_is_noop():
    from llamax_b200.dp import GradBucket

    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    GradBucket([p]).allreduce_()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
