"""world_size-2 gloo tests (CPU) of the N>1 host logic: sample partition, round-robin tensor sharding
with the scalar-KL all-reduce, and the flat-gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import variational_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bayesianneuralnetworks_b200 import parallel
        # --- sample partition: the union over ranks is [0, S), blocks are contiguous and equal
        b, e = parallel.sample_range(8, rank, world)
        got = [None] * world
        dist.all_gather_object(got, (b, e))
        assert sorted(got) == [(0, 4), (4, 8)]
        assert parallel.grid_coordinates(rank, world, 2) == (0, rank)
        # --- the data x sample grid bench.py / training.ElboTrainer use by default: 2 = 1 x 2, 4 = 2 x 2, 8 = 2 x 4;
        #     every rank of a data group sees the same batch slice, the sample blocks of a data group tile [0, S)
        assert [parallel.default_grid(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
        for w in (2, 4, 8):
            dg, sg = parallel.default_grid(w)
            cells = [parallel.grid_coordinates(r, w, sg) for r in range(w)]
            assert sorted(cells) == [(d, s) for d in range(dg) for s in range(sg)]
            for d in range(dg):
                blocks = sorted(parallel.sample_range(16 * sg, s, sg) for dd, s in cells if dd == d)
                assert blocks[0][0] == 0 and blocks[-1][1] == 16 * sg
                assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        # --- tensor-sharded KL with one small all-reduce == the oracle's mean-of-means
        g = torch.Generator().manual_seed(0)
        tensors = [(torch.rand(n, generator=g) - 0.5, torch.randn(n, generator=g) * 0.15 - 2.0, 0.0, 0.1)
                   for n in (1000, 10, 333, 7, 64)]
        mine = parallel.round_robin(list(range(len(tensors))), rank, world)
        sums = torch.zeros(len(tensors), dtype=torch.float64)
        for i in mine:
            sums[i] = orc.kl_tensor_sums([tensors[i]])[0]
        total = parallel.allreduce_kl_sums(sums, None, [t[0].numel() for t in tensors], 7)
        want = orc.kl_divergence(tensors, 7)
        assert float(total) == pytest.approx(float(want), rel=1e-6)
        # --- flat gradient all-reduce: average of per-rank gradients, missing grads treated as zeros
        torch.manual_seed(1)
        lin = torch.nn.Linear(5, 3)
        extra = torch.nn.Parameter(torch.ones(4))
        x = torch.full((2, 5), float(rank + 1))
        lin(x).sum().backward()
        if rank == 0:
            (extra * 2).sum().backward()
        parallel.allreduce_gradients(list(lin.parameters()) + [extra])
        assert torch.allclose(lin.weight.grad, torch.full((3, 5), 3.0))       # (2*1 + 2*2) / 2
        assert torch.allclose(lin.bias.grad, torch.full((3,), 2.0))
        assert torch.allclose(extra.grad, torch.full((4,), 1.0))              # (2 + 0) / 2
        # --- parameters kept in channels_last (bench.py's torch trunk): gradients are reduced in logical order
        conv = torch.nn.Conv2d(2, 3, 3).to(memory_format=torch.channels_last)
        conv(torch.full((1, 2, 4, 4), float(rank + 1))).sum().backward()
        parallel.allreduce_gradients(list(conv.parameters()))
        assert torch.allclose(conv.weight.grad, torch.full((3, 2, 3, 3), 6.0))   # 4 positions * (1 + 2) / 2
        assert torch.allclose(conv.bias.grad, torch.full((3,), 4.0))
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}
