"""Batches straight from a dataset's memory maps (reference: permutect/data/memory_mapped_data.py:36-58 and
permutect/data/reads_dataset.py:109-196).

The reference walks the maps one Datum at a time (a Python loop builds ``read_end_indices``; ``__iter__`` yields Datum
objects; ``Batch.__init__`` re-stacks ref and alt rows with one ``np.vstack`` per field): ~6 k variants/s per worker,
slower than the model on a CPU.  Here a batch is three contiguous slices -- rows of the int16 map, rows of the fp16
map, and the matching slice of the compressed-reads map in DATASET order -- and the ref/alt regrouping becomes a
gather-index array written on the device (``Batch.from_dataset_slice`` -> pmt_dataset_read_indices).
"""
from typing import Iterator, Optional

import numpy as np

from permutect_b200.data.batch import Batch
from permutect_b200.data.datum import Data


class MemoryMappedBatches:
    """Iterable of ``Batch`` over consecutive variants of (int_mmap [N, 16+2L] int16, float_mmap [N, 6+I] fp16,
    reads_mmap [R, row_bytes] uint8).  ``num_data`` / ``num_reads`` bound the valid prefix of the maps, as in the
    reference (the files may be larger than the data, memory_mapped_data.py:45-47).  ``start`` / ``stop`` select a variant
    range (the contiguous shard of one rank or worker, reads_dataset.py:141-142)."""

    def __init__(self, int_mmap, float_mmap, reads_mmap, batch_size: int, num_data: Optional[int] = None,
                 start: int = 0, stop: Optional[int] = None, pin_memory: bool = True):
        self.int_mmap, self.float_mmap, self.reads_mmap = int_mmap, float_mmap, reads_mmap
        self.num_data = len(int_mmap) if num_data is None else num_data
        self.batch_size = int(batch_size)
        self.start, self.stop = start, self.num_data if stop is None else min(stop, self.num_data)
        self.pin_memory = pin_memory
        counts = np.asarray(int_mmap[: self.num_data, : Data.ALT_COUNT.idx + 1]).astype(np.int64)
        # read_end_indices (memory_mapped_data.py:53-58) as one cumulative sum; [v] = first row of variant v
        self.read_start_indices = np.concatenate(([0], np.cumsum(counts[:, Data.REF_COUNT.idx] + counts[:, Data.ALT_COUNT.idx])))

    def __len__(self) -> int:
        return (self.stop - self.start + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Batch]:
        for v0 in range(self.start, self.stop, self.batch_size):
            v1 = min(v0 + self.batch_size, self.stop)
            r0, r1 = int(self.read_start_indices[v0]), int(self.read_start_indices[v1])
            batch = Batch.from_dataset_slice(np.asarray(self.int_mmap[v0:v1]), np.asarray(self.float_mmap[v0:v1]),
                                             np.asarray(self.reads_mmap[r0:r1]))
            yield batch.pin_memory() if self.pin_memory else batch
