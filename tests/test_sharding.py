"""N>1 host logic on CPU: contiguous sharding + the all-gather of per-instance results, with
two gloo ranks. Each rank solves its shard with the oracle (the GPU is not needed to test the
plumbing) and the gathered result must equal a single-process solve of the whole batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from path_optimizer_2_b200 import abi, sharding, synthetic


def test_shard_range_partitions():
    for total in (1, 7, 8, 1024, 65536):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    cost = torch.tensor([1.5, -2.0, 3.25], dtype=torch.float64)
    status = torch.tensor([0, 1, 2], dtype=torch.int32)
    iters = torch.tensor([25, 4000, 50], dtype=torch.int32)
    c, s, i = sharding.unpack_results(sharding.pack_results(cost, status, iters))
    assert torch.equal(c, cost) and torch.equal(s, status) and torch.equal(i, iters)


def _worker(rank, world, port, total, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    lo, hi = sharding.shard_range(total, rank, world)
    hb = synthetic.make_batch(3, hi - lo, n, first=lo)
    res, _ = oracle.solve_batch(abi.default_params(), hb, nthreads=1)
    packed = sharding.pack_results(torch.from_numpy(res.cost), torch.from_numpy(res.status),
                                   torch.from_numpy(res.iters))
    allr = sharding.gather_results(packed)
    c, s, i = sharding.unpack_results(allr)
    if rank == 0:
        q.put((c.numpy(), s.numpy(), i.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    from oracle import oracle
    total, n, world = 8, 40, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    c, s, i = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, _ = oracle.solve_batch(abi.default_params(), synthetic.make_batch(3, total, n), nthreads=1)
    assert np.array_equal(c, ref.cost) and np.array_equal(s, ref.status) and np.array_equal(i, ref.iters)
