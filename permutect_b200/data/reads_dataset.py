"""Batches straight from a dataset's memory maps (reference: permutect/data/memory_mapped_data.py:36-58 and
permutect/data/reads_dataset.py:109-196).

The reference walks the maps one Datum at a time (a Python loop builds ``read_end_indices``; ``__iter__`` yields Datum
objects; ``Batch.__init__`` re-stacks ref and alt rows with one ``np.vstack`` per field): ~6 k variants/s per worker,
slower than the model on a CPU.  Here a batch is three contiguous slices -- rows of the int16 map, rows of the fp16
map, and the matching slice of the compressed-reads map in DATASET order -- and the ref/alt regrouping becomes a
gather-index array written on the device (``Batch.from_dataset_slice`` -> pmt_dataset_read_indices).
"""
import random
from typing import Iterator, List, Optional

import numpy as np
import torch

from permutect_b200.data.batch import Batch, BatchIndexedTensor, BatchProperty
from permutect_b200.data.datum import Data, Datum, HAPLOTYPES_START_IDX, INFO_START_IDX, num_read_features


class _StagingSlot:
    """Persistent pinned host buffers of one in-flight batch (int / float / reads), grown on demand."""

    def __init__(self):
        self.buffers = {}
        self.last_batch = None      # the Batch last built over these buffers (its H2D event tells when they are free again)

    def view(self, name: str, like: np.ndarray) -> np.ndarray:
        buf = self.buffers.get(name)
        nbytes = like.size * like.dtype.itemsize
        if buf is None or buf.numel() < nbytes:
            buf = self.buffers[name] = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, pin_memory=True)
        return buf[:nbytes].numpy().view(like.dtype).reshape(like.shape)


def _register_host_array(arr):
    """Page-locks the memory of a host array in place (cudaHostRegister) so that batches cut from it are DMA-able views: no
    staging copy at all.  Works for arrays that live in anonymous memory (a dataset loaded into RAM); a file-backed map is
    usually refused by the driver -- the caller then falls back to the staging ring.  Returns None on failure, else the
    pointer to hand to ``_unregister_host_arrays`` (0 when the memory was pinned already and is not ours to release).  The
    registration MUST end before the memory is freed: a freed range that still counts as page-locked makes later host
    tensors at the same addresses look pinned, and their non-blocking copies race with the host."""
    if not isinstance(arr, np.ndarray) or not arr.flags["C_CONTIGUOUS"] or arr.nbytes == 0:
        return None
    try:
        if torch.from_numpy(arr[:1]).is_pinned():
            return 0
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
        if int(rc) == 0 and torch.from_numpy(arr[:1]).is_pinned():
            return int(arr.ctypes.data)
        return None
    except Exception:
        return None


def _unregister_host_arrays(pointers):
    for ptr in pointers:
        if ptr:
            try:
                torch.cuda.cudart().cudaHostUnregister(ptr)
            except Exception:
                pass


def load_into_pinned_memory(array) -> np.ndarray:
    """A copy of a dataset array in page-locked host memory obtained from the CUDA allocator (``cudaHostAlloc``), as a numpy
    array: what to load a dataset's maps into ONCE when the whole dataset fits in RAM.  Batches cut from such arrays are
    DMA-able views like those of arrays registered in place, and under the load of eight ranks sharing one host this memory
    is served faster than memory registered after the fact."""
    src = np.asarray(array)
    t = torch.empty(src.shape, dtype=torch.from_numpy(src[:0]).dtype, pin_memory=True)
    out = t.numpy()              # keeps the tensor (and its allocation) alive through .base
    np.copyto(out, src)
    return out


def _parallel_copy(pool, dst: np.ndarray, src: np.ndarray, n_parts: int):
    """dst[:] = src in row ranges handed to the pool (numpy releases the GIL inside the copy loop, so the ranges are copied
    by different cores: one core moves ~8 GB/s out of the page cache, an ingest pipeline needs a multiple of that)."""
    n = len(src)
    if n_parts <= 1 or n < 4096:
        np.copyto(dst, src)
        return []
    step = (n + n_parts - 1) // n_parts
    return [pool.submit(np.copyto, dst[a:a + step], src[a:a + step]) for a in range(0, n, step)]


class MemoryMappedBatches:
    """Iterable of ``Batch`` over consecutive variants of (int_mmap [N, 16+2L] int16, float_mmap [N, 6+I] fp16,
    reads_mmap [R, row_bytes] uint8).  ``num_data`` / ``num_reads`` bound the valid prefix of the maps, as in the
    reference (the files may be larger than the data, memory_mapped_data.py:45-47).  ``start`` / ``stop`` select a variant
    range (the contiguous shard of one rank or worker, reads_dataset.py:141-142).

    ``staging_threads > 0`` (with ``pin_memory``): a producer thread cuts the batches ``prefetch`` ahead of the consumer
    and copies their three slices into a ring of persistent pinned buffers with that many copy threads -- what the
    reference leaves to DataLoader worker processes (reads_dataset.py:223-232), without the per-Datum collate.  The ring
    holds ``prefetch + 3`` batches: a batch's buffers are refilled once that many further batches have been cut AND the
    host-to-device copy ``prefetch_generator`` made of it has finished (its event is waited for); a consumer that does not go
    through ``prefetch_generator`` must be done with a batch before it has taken ``prefetch + 2`` more.

    ``pin_memory="register"``: the three arrays are page-locked IN PLACE once (cudaHostRegister) and every batch is a
    zero-copy view of them -- the host side of the ingest pipeline is then three slices per batch.  For arrays the driver
    refuses to register (file-backed maps) this falls back to the staging ring (``staging_threads`` or 4 copy threads)."""

    def __init__(self, int_mmap, float_mmap, reads_mmap, batch_size: int, num_data: Optional[int] = None,
                 start: int = 0, stop: Optional[int] = None, pin_memory: bool = True, staging_threads: int = 0, prefetch: int = 2):
        self.int_mmap, self.float_mmap, self.reads_mmap = int_mmap, float_mmap, reads_mmap
        self.num_data = len(int_mmap) if num_data is None else num_data
        self.batch_size = int(batch_size)
        self.start, self.stop = start, self.num_data if stop is None else min(stop, self.num_data)
        self.staging_threads, self.prefetch = int(staging_threads), max(1, int(prefetch))
        self._slots = None
        self.registered = False
        if pin_memory == "register":
            arrays = [np.asarray(a) if not isinstance(a, np.memmap) else a for a in (int_mmap, float_mmap, reads_mmap)]
            handles = []
            for arr in arrays:
                h = _register_host_array(arr)
                if h is None:
                    break
                handles.append(h)
            self.registered = len(handles) == len(arrays)
            if self.registered:
                import weakref
                self.int_mmap, self.float_mmap, self.reads_mmap = arrays       # the loader keeps the page-locked arrays alive ...
                self._unregister = weakref.finalize(self, _unregister_host_arrays, handles)   # ... and unlocks them when it goes
            else:
                _unregister_host_arrays(handles)
                self.staging_threads = self.staging_threads or 4
            pin_memory = True
        self.pin_memory = bool(pin_memory)
        counts = np.asarray(int_mmap[: self.num_data, : Data.ALT_COUNT.idx + 1]).astype(np.int64)
        # read_end_indices (memory_mapped_data.py:53-58) as one cumulative sum; [v] = first row of variant v
        self.read_start_indices = np.concatenate(([0], np.cumsum(counts[:, Data.REF_COUNT.idx] + counts[:, Data.ALT_COUNT.idx])))

    def __len__(self) -> int:
        return (self.stop - self.start + self.batch_size - 1) // self.batch_size

    def _slices(self):
        for v0 in range(self.start, self.stop, self.batch_size):
            v1 = min(v0 + self.batch_size, self.stop)
            r0, r1 = int(self.read_start_indices[v0]), int(self.read_start_indices[v1])
            yield self.int_mmap[v0:v1], self.float_mmap[v0:v1], self.reads_mmap[r0:r1]

    def __iter__(self) -> Iterator[Batch]:
        if self.registered:
            for ints, floats, reads in self._slices():
                yield Batch.from_dataset_slice(ints, floats, reads)      # views of page-locked memory
            return
        if self.pin_memory and self.staging_threads > 0:
            yield from self._staged()
            return
        for ints, floats, reads in self._slices():
            batch = Batch.from_dataset_slice(np.asarray(ints), np.asarray(floats), np.asarray(reads))
            yield batch.pin_memory() if self.pin_memory else batch

    def _staged(self) -> Iterator[Batch]:
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        if self._slots is None:
            self._slots = [_StagingSlot() for _ in range(self.prefetch + 3)]
        slots, out = self._slots, queue.Queue(maxsize=self.prefetch)
        stop_flag = threading.Event()

        def produce():
            try:
                with ThreadPoolExecutor(self.staging_threads) as pool:
                    for i, (ints, floats, reads) in enumerate(self._slices()):
                        if stop_flag.is_set():
                            return
                        slot = slots[i % len(slots)]
                        done = getattr(slot.last_batch, "_h2d_done", None) if slot.last_batch is not None else None
                        if done is not None:
                            done.synchronize()          # the copy out of these buffers has left the host
                        views = [slot.view(name, np.asarray(src)) for name, src in (("int", ints), ("float", floats), ("reads", reads))]
                        jobs = []
                        for dst, src in zip(views, (ints, floats, reads)):
                            jobs += _parallel_copy(pool, dst, np.asarray(src), self.staging_threads)
                        for j in jobs:
                            j.result()
                        batch = Batch.from_dataset_slice(*views)
                        slot.last_batch = batch
                        out.put(batch)
                out.put(None)
            except BaseException as e:      # hand the failure to the consumer instead of dying silently
                out.put(e)

        worker = threading.Thread(target=produce, name="pmt-staging", daemon=True)
        worker.start()
        try:
            while True:
                item = out.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop_flag.set()
            while worker.is_alive():        # unblock a producer waiting on the full queue
                try:
                    out.get_nowait()
                except queue.Empty:
                    worker.join(timeout=0.05)


# ---- fold helpers (reads_dataset.py:26-40) ---------------------------------------------------------------------------
def last_fold_only(num_folds: int):
    return [num_folds - 1]


def all_but_last_fold(num_folds: int):
    return list(range(num_folds - 1))


def all_but_one_fold(num_folds: int, fold_to_exclude: int):
    return list(range(fold_to_exclude)) + list(range(fold_to_exclude + 1, num_folds))


def all_folds(num_folds: int):
    return list(range(num_folds))


def batch_order_rows(read_start: np.ndarray, ref_counts: np.ndarray, alt_counts: np.ndarray, order: np.ndarray) -> np.ndarray:
    """Row indices into a dataset-order reads array that lay the variants ``order`` out as a batch: all their ref rows, then
    all their alt rows (batch.py:45-47).  ``read_start[v]`` is the first row of variant v."""
    rc, ac = ref_counts[order], alt_counts[order]
    n_ref, n_alt = int(rc.sum()), int(ac.sum())
    ref_first = np.concatenate(([0], np.cumsum(rc)))[:-1]
    alt_first = np.concatenate(([0], np.cumsum(ac)))[:-1]
    ref_rows = np.repeat(read_start[order] - ref_first, rc) + np.arange(n_ref)
    alt_rows = np.repeat(read_start[order] + rc - alt_first, ac) + np.arange(n_alt)
    return np.concatenate((ref_rows, alt_rows))


class ReadsDataset:
    """reads_dataset.py:46-196 over a ``MemoryMappedData``: fold selection, totals by source / label / variant type / counts,
    and the reference's iteration order -- the variant range is cut into chunks that fit in RAM, the chunks are visited in
    shuffled order and the variants of a chunk in shuffled order (``random.shuffle``, so a seeded ``random`` reproduces the
    reference's order exactly).  ``__iter__`` yields Datum objects like the reference; ``batches()`` / ``make_data_loader()``
    yield the same variants in the same order already collated: one fancy-index gather per array and batch instead of one
    Python object per variant."""

    def __init__(self, memory_mapped_data, num_folds: int = 1, folds_to_use: Optional[List[int]] = None, keep_probs_by_label_l=None,
                 chunks: Optional[int] = None):
        self.memory_mapped_data = memory_mapped_data.restrict_to_folds(num_folds, folds_to_use, keep_probs_by_label_l)
        d = self.memory_mapped_data
        self._size = d.num_data
        self._read_end_indices = d.read_end_indices
        self._int_array_ve, self._float_array_ve, self._stacked_reads_re = d.int_mmap, d.float_mmap, d.reads_mmap
        self._chunks = chunks
        ints = np.asarray(d.int_mmap[: d.num_data])
        self._num_read_features = num_read_features(d.reads_mmap.shape[1]) if d.reads_mmap is not None else 0
        self._num_info_features = d.float_mmap.shape[1] - INFO_START_IDX
        self._haplotypes_length = d.int_mmap.shape[1] - HAPLOTYPES_START_IDX
        # totals_slvra (reads_dataset.py:83-88) without a Datum loop: one scatter-add over the whole int array
        n_sources = int(ints[:, Data.SOURCE.idx].max()) + 1 if len(ints) else 1
        self.totals_slvra = BatchIndexedTensor.zeros(num_sources=n_sources, include_logits=False, device=torch.device("cpu"))
        if len(ints):
            probe = Batch.__new__(Batch)
            probe.int_tensor = torch.from_numpy(np.ascontiguousarray(ints))
            probe._size = len(ints)
            probe._device_counts = None
            probe.lazy_batch_indices = {False: None, True: None}
            self.totals_slvra.record(probe, torch.ones(len(ints)))
        self.totals_by_label_l = self.totals_slvra.get_marginal(BatchProperty.LABEL)

    def __len__(self):
        return self._size

    def totals_by_label(self):
        return self.totals_by_label_l

    def num_read_features(self) -> int:
        return self._num_read_features

    def num_info_features(self) -> int:
        return self._num_info_features

    def haplotypes_length(self) -> int:
        return self._haplotypes_length

    def num_sources(self) -> int:
        return self.totals_slvra.num_sources()

    def _chunk_count(self) -> int:
        if self._chunks is not None:
            return self._chunks
        import psutil
        return 1 + ((8 * self.memory_mapped_data.num_bytes()) // psutil.virtual_memory().available)   # reads_dataset.py:126-128

    def _chunk_plan(self, start: int, stop: int):
        """(chunk start, chunk stop, shuffled local order) in the reference's visiting order (reads_dataset.py:141-176)."""
        n_chunks = self._chunk_count()
        per_chunk = (stop - start) // n_chunks
        chunks = list(range(n_chunks))
        random.shuffle(chunks)
        for chunk in chunks:
            c0 = start + chunk * per_chunk
            c1 = start + (chunk + 1) * per_chunk if chunk < n_chunks - 1 else stop
            indices = list(range(c1 - c0))
            random.shuffle(indices)
            yield c0, c1, np.asarray(indices, dtype=np.int64)

    def __iter__(self) -> Iterator[Datum]:
        d = self.memory_mapped_data
        for c0, c1, order in self._chunk_plan(0, self._size):
            for idx in order + c0:
                r0, r1 = int(d.read_start_indices[idx]), int(d.read_start_indices[idx + 1])
                yield Datum(d.int_mmap[idx], d.float_mmap[idx], d.reads_mmap[r0:r1], compressed=True)

    def batches(self, batch_size: int, pin_memory: bool = False, start: int = 0, stop: Optional[int] = None) -> Iterator[Batch]:
        """The batches ``DataLoader(self, batch_size, collate_fn=Batch)`` builds from ``__iter__`` (make_data_loader,
        reads_dataset.py:200-209), for variants [start, stop) -- the contiguous shard of one rank."""
        d = self.memory_mapped_data
        stop = self._size if stop is None else min(stop, self._size)
        carry = None       # variants of a chunk's tail wait for the next chunk (a DataLoader batch may straddle chunks)
        for c0, c1, order in self._chunk_plan(start, stop):
            r0, r1 = int(d.read_start_indices[c0]), int(d.read_start_indices[c1])
            ints, floats = np.asarray(d.int_mmap[c0:c1]), np.asarray(d.float_mmap[c0:c1])     # the chunk, sequentially, into RAM
            reads = np.asarray(d.reads_mmap[r0:r1])
            rstart = d.read_start_indices[c0:c1] - r0
            rc = ints[:, Data.REF_COUNT.idx].astype(np.int64)
            ac = ints[:, Data.ALT_COUNT.idx].astype(np.int64)
            pos = 0
            if carry is not None:
                take = min(batch_size - len(carry[0]), len(order))
                sel = order[:take]
                pos = take
                part = (ints[sel], floats[sel], [reads[rstart[v]:rstart[v] + rc[v]] for v in sel], [reads[rstart[v] + rc[v]:rstart[v] + rc[v] + ac[v]] for v in sel])
                carry = (np.concatenate((carry[0], part[0])), np.concatenate((carry[1], part[1])), carry[2] + part[2], carry[3] + part[3])
                if len(carry[0]) == batch_size:
                    yield self._finish(Batch.from_arrays(carry[0], carry[1], np.concatenate(carry[2] + carry[3])), pin_memory)
                    carry = None
            while pos + batch_size <= len(order):
                sel = order[pos:pos + batch_size]
                pos += batch_size
                yield self._finish(Batch.from_arrays(ints[sel], floats[sel], reads[batch_order_rows(rstart, rc, ac, sel)]), pin_memory)
            if pos < len(order):
                sel = order[pos:]
                carry = (ints[sel], floats[sel], [reads[rstart[v]:rstart[v] + rc[v]] for v in sel],
                         [reads[rstart[v] + rc[v]:rstart[v] + rc[v] + ac[v]] for v in sel])
        if carry is not None:
            yield self._finish(Batch.from_arrays(carry[0], carry[1], np.concatenate(carry[2] + carry[3])), pin_memory)

    @staticmethod
    def _finish(batch: Batch, pin_memory: bool) -> Batch:
        return batch.pin_memory() if pin_memory else batch

    def make_data_loader(self, batch_size: int, pin_memory: bool = False, num_workers: int = 0):
        """An iterable of Batch (a fresh shuffled pass per ``iter()``).  ``num_workers`` is accepted for API parity: collation
        is a handful of numpy gathers per batch, there is nothing left to hand to worker processes."""
        dataset = self

        class _Loader:
            def __iter__(self_inner):
                return dataset.batches(batch_size, pin_memory)

            def __len__(self_inner):
                return (len(dataset) + batch_size - 1) // batch_size
        return _Loader()
