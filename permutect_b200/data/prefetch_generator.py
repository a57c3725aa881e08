"""Host -> device batch pipeline (reference: permutect/data/prefetch_generator.py:9-20, whose overlapped variant is
commented out at :22-36).  Batches produced by the loader (pinned host memory, reads still compressed: 12 B/read)
are copied on a side CUDA stream into a ring of persistent device buffers, up to ``depth`` batches ahead of the batch
the model is working on, so PCIe transfers overlap the kernels.  The consumer's stream waits on the copy's event and
the copy stream waits on the event that marks the consumer's last use of a ring slot; there is no host
synchronisation and, after the first pass, no device allocation.  The ring (buffers, copy stream, release events)
outlives a generator: the next pass over a loader (the next epoch, the next evaluation pass) starts copying while the
kernels of the previous pass are still running instead of waiting for the compute stream to drain."""
import copy
from typing import Iterable, Iterator, List, Optional

import torch

from permutect_b200.data.batch import Batch

_FIELDS = ("reads", "int_tensor", "float_tensor", "read_indices")


class _Slot:
    """One ring entry: flat device buffers per batch field, grown on demand, plus the event of its last consumer."""

    def __init__(self, device):
        self.device = device
        self.buffers = {}
        self.released: Optional[torch.cuda.Event] = None
        self.fresh = False    # a buffer was (re)allocated since the last copy: its memory may still be in use on the caller's stream

    def view_like(self, name: str, t_cpu: torch.Tensor) -> torch.Tensor:
        buf = self.buffers.get(name)
        n = t_cpu.numel()
        if buf is None or buf.dtype != t_cpu.dtype or buf.numel() < n:
            with torch.inference_mode(False):     # the ring outlives this pass: an inference tensor could not be refilled by a training pass
                buf = torch.empty(int(n * 1.1) + 64, dtype=t_cpu.dtype, device=self.device)
            self.buffers[name] = buf
            self.fresh = True
        return buf[:n].view(t_cpu.shape)


class _Ring:
    def __init__(self, device, depth: int):
        self.copy_stream = torch.cuda.Stream(device)
        self.slots: List[_Slot] = [_Slot(device) for _ in range(depth + 1)]
        self.in_use = False


_RINGS = {}   # (device, depth) -> _Ring kept between generators


def prefetch_generator(dataloader: Iterable[Batch], device=None, depth: int = 2) -> Iterator[Batch]:
    """LIFETIME CONTRACT: a yielded batch's tensors are views into a ring of ``depth + 1`` persistent device buffers; the
    slot is refilled (on the copy stream, ordered after everything the consumer queued on its CURRENT stream up to the moment
    it asks for the next batch) once ``depth`` further batches have been requested.  A consumer that keeps a batch, or
    anything derived from its int / float / reads views, beyond its own iteration -- ``list(prefetch_generator(...))``, a
    recorder holding references, work queued on another stream -- must call ``batch.detach_from_ring()`` (device-side
    clone) first.  The reference's generator yields independent tensors (prefetch_generator.py:9-20)."""
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("prefetch_generator feeds the CUDA kernels: there is no CPU path (device must be a CUDA device)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    shared = _RINGS.get((device, depth))
    if shared is None:
        shared = _RINGS[(device, depth)] = _Ring(device, depth)
    owner = not shared.in_use          # two generators alive at once (nested loaders) must not share slots
    holder = shared if owner else _Ring(device, depth)
    holder.in_use = True
    try:
        yield from _pipeline(dataloader, device, depth, holder)
    finally:
        holder.in_use = False


def _pipeline(dataloader: Iterable[Batch], device, depth: int, holder: "_Ring") -> Iterator[Batch]:
    copy_stream, ring = holder.copy_stream, holder.slots
    queue = []
    n_launched = 0

    def launch(batch_cpu: Batch):
        nonlocal n_launched
        slot = ring[n_launched % len(ring)]
        n_launched += 1
        views = {name: slot.view_like(name, getattr(batch_cpu, name)) for name in _FIELDS if getattr(batch_cpu, name) is not None}
        if slot.released is not None:
            copy_stream.wait_event(slot.released)       # the consumer of this slot's previous batch is done with it
        if slot.fresh or slot.released is None:
            copy_stream.wait_stream(torch.cuda.current_stream(device))   # buffers were just allocated on the caller's stream
            slot.fresh = False
        with torch.cuda.stream(copy_stream):
            for name, dst in views.items():
                dst.copy_(getattr(batch_cpu, name), non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        batch_cpu._h2d_done = done          # a host-side staging ring (reads_dataset.MemoryMappedBatches) reuses the buffers after this
        batch_gpu = copy.copy(batch_cpu)
        for name, dst in views.items():
            setattr(batch_gpu, name, dst)
        batch_gpu._offsets = None
        batch_gpu._decoded = None
        batch_gpu.lazy_batch_indices = {False: None, True: None}
        if getattr(batch_cpu, "_dataset_order", False):
            batch_gpu.read_indices = None
        queue.append((batch_gpu, done, slot))

    def hand_over():
        batch_gpu, done, slot = queue.pop(0)
        torch.cuda.current_stream(device).wait_event(done)
        batch_gpu.finalize_on_device()      # e.g. gather indices of a dataset-order batch, on the consumer's stream
        return batch_gpu, slot

    def release(slot: _Slot):
        slot.released = torch.cuda.Event()
        slot.released.record(torch.cuda.current_stream(device))   # everything the consumer enqueued on this batch

    for batch_cpu in dataloader:
        launch(batch_cpu)
        if len(queue) >= depth:
            batch_gpu, slot = hand_over()
            try:
                yield batch_gpu
            finally:
                release(slot)                  # also when the consumer stops early: the ring outlives this generator
    while queue:
        batch_gpu, slot = hand_over()
        try:
            yield batch_gpu
        finally:
            release(slot)
