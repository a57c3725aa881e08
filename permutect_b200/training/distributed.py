"""Variant-sharded data parallelism.  The reference has no multi-GPU code; variants are
independent (the only cross-read coupling is inside a variant, gated_mlp.py:236-248), so

  * inference shards variants by contiguous index range with NO collective (the scheme the reference uses
    for DataLoader workers, reads_dataset.py:141-142);
  * training adds exactly one exchange per optimiser step: an all-reduce(SUM) of the flat fp32 gradient
    (68 229 floats = 273 KB at the shipped hyper-parameters).  The reference differentiates the SUM of the
    per-variant losses (artifact_model.py:90), so gradients are summed, not averaged, and the learning rate
    is unchanged; clipping uses the norm of the reduced gradient, hence identical updates on every rank.
    Host-side per-rank statistics that feed back into training (Balancer counters) are summed the same way.
"""
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) of rank's share of n_items (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_flat_(tensors: List[torch.Tensor], group=None) -> None:
    """Sum a list of same-dtype tensors across ranks with ONE collective on a flat buffer, in place."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> None:
    """Sum .grad across ranks.  A parameter without a gradient on this rank contributes zeros, so the
    collective has the same shape everywhere (e.g. calibration epochs freeze most tensors on all ranks alike)."""
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    allreduce_flat_([p.grad for p in params], group)


def allreduce_counters(tensors: List[torch.Tensor], group=None) -> None:
    """Balancer / loss-metric increments observed on each rank -> global sums (balancer.py:61-72)."""
    allreduce_flat_(tensors, group)


class SyncedBalancer:
    """Keeps a ``Balancer``'s stratified counters global under data parallelism (balancer.py:57-72).

    Before the wrapped balancer sees a rank's batch, the increments every rank is about to make -- one per variant into
    ``counts_slvra``, the artifact / non-artifact probability mass of unlabeled variants into ``pseudo_counts_slvra``, the
    batch size into ``count_since_last_recomputation`` -- are summed with ONE all-reduce, and the other ranks' share is added
    to the wrapped balancer's state.  The balancer then processes its own batch exactly as in a single process, so its
    periodic weight recomputation (balancer.py:74-104) fires on the same step and from the same totals on every rank.
    Everything else (``weights_slvra``, plots, ``state_dict``) is the wrapped object's: attribute access falls through."""

    def __init__(self, balancer, group=None):
        self.balancer, self.group = balancer, group

    def __getattr__(self, name):
        return getattr(self.balancer, name)

    def process_batch_and_compute_weights(self, batch, artifact_probs_b: torch.Tensor):
        b = self.balancer
        counts, pseudo = b.counts_slvra, b.pseudo_counts_slvra
        dev = counts.device
        idx = batch.batch_indices()
        probs = artifact_probs_b.to(dev)
        unlabeled = (1 - batch.get_is_labeled_mask()).to(dev)
        n = counts.numel()
        local = torch.zeros(2 * n + 1, dtype=counts.dtype, device=dev)
        flat_idx = idx.flattened_idx.to(dev)
        stride = counts[0, 0].numel()         # one label to the next
        local[:n].index_add_(0, flat_idx, torch.ones(batch.size(), dtype=counts.dtype, device=dev))
        base = flat_idx - stride * idx.labels.to(dev)        # label axis zeroed; ARTIFACT = 0, VARIANT = 1 (utils/enums.py)
        local[n:2 * n].index_add_(0, base, (unlabeled * probs).to(counts.dtype))
        local[n:2 * n].index_add_(0, base + stride, (unlabeled * (1 - probs)).to(counts.dtype))
        local[2 * n] = batch.size()
        total = local.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        others = total - local
        with torch.no_grad():
            counts.view(-1).add_(others[:n])
            pseudo.view(-1).add_(others[n:2 * n])
        b.count_since_last_recomputation += int(round(float(others[2 * n])))
        return b.process_batch_and_compute_weights(batch, artifact_probs_b)


def gather_variant_outputs(local: torch.Tensor, n_total: int, group=None) -> Optional[torch.Tensor]:
    """Concatenate rank-local per-variant outputs in variant order on every rank (inference needs the order
    only to re-associate logits with their Datum, filter_variants.py:302-320)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank]
    biggest = max(sizes)          # shard sizes differ by at most one; pad so the collective is uniform
    padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    chunks = [torch.empty_like(padded) for _ in sizes]
    dist.all_gather(chunks, padded, group=group)
    return torch.cat([c[:s] for c, s in zip(chunks, sizes)])


def bind_to_gpu_numa_node(local_rank: int) -> Optional[List[int]]:
    """Pins the calling process to the CPU cores NVML reports as local to GPU ``local_rank`` (one process per GPU: pinned
    host buffers are then allocated on, and the H2D / D2H copies of the ingest pipeline stay on, the socket the GPU hangs
    off instead of crossing the inter-socket link).  Returns the core list, or None when NVML / affinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
