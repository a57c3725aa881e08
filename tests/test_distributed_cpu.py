"""world_size-2 gloo tests of the data-parallel host logic (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from permutect_b200.training import distributed as pdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [pdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from golden_utils import load
        from helpers import model_from_golden
        g = load("small_hp")
        model = model_from_golden(g, torch.device("cpu"))
        # every rank holds the gradient of its own shard; here: golden gradient scaled by (rank + 1)
        for name, p in model.named_parameters():
            p.grad = torch.from_numpy(g.grad[name]).clone() * (rank + 1)
        frozen = model.feature_clustering.artifact_emg.mu_k
        if rank == 1:
            frozen.grad = None                       # missing on one rank: must contribute zeros, not hang
        pdist.allreduce_gradients(model.parameters())
        ok = True
        for name, p in model.named_parameters():
            want = torch.from_numpy(g.grad[name]) * (3.0 if p is not frozen else 1.0)
            ok = ok and torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7)
        counters = [torch.full((3, 5), float(rank + 1)), torch.arange(4.0) * (rank + 1)]
        pdist.allreduce_counters(counters)
        ok = ok and torch.equal(counters[0], torch.full((3, 5), 3.0)) and torch.equal(counters[1], torch.arange(4.0) * 3)
        n_total = 11
        a, b = pdist.shard_range(n_total, rank, world)
        gathered = pdist.gather_variant_outputs(torch.arange(a, b, dtype=torch.float32)[:, None], n_total)
        ok = ok and torch.equal(gathered[:, 0], torch.arange(n_total, dtype=torch.float32))
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_numa_binding_helper_is_harmless_without_a_gpu():
    """bind_to_gpu_numa_node: the core list NVML reports for the GPU, or None (no NVML / no such GPU) -- never an exception,
    and the process keeps a non-empty affinity mask."""
    import os
    from permutect_b200.training.distributed import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    cores = bind_to_gpu_numa_node(0)
    assert cores is None or (len(cores) > 0 and set(cores) <= before)
    assert len(os.sched_getaffinity(0)) > 0
    os.sched_setaffinity(0, before)
