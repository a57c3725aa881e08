"""The dataset container of the reference (permutect/data/memory_mapped_data.py:35-285) in the same on-disk format: a tar of

    metadata.metadata.npy         torch.save of uint32 [num_data, int columns, float columns, num_reads, read row bytes]
    int_array.int_mmap.npy        int16  [capacity >= num_data, 16 + 2L]   (datum.py:51-81)
    float_array.float_mmap.npy    fp16   [capacity >= num_data, 6 + I]
    reads_array.reads_mmap.npy    uint8  [capacity >= num_reads, row bytes]  variant after variant: ref rows, then alt rows

so a dataset written by the reference's preprocessing loads here and vice versa.  Everything the reference does with a
Python loop over Datum objects is array arithmetic: ``read_end_indices`` is one cumulative sum
(memory_mapped_data.py:53-58), fold selection a boolean mask over ``index % num_folds`` (:85-101).
"""
import os
import random
import tarfile
import tempfile
from typing import Generator, List, Optional

import numpy as np
import torch

from permutect_b200.data.datum import COMPRESSED_READS_ARRAY_DTYPE, Data, Datum

SUFFIX_FOR_INT_MMAP = ".int_mmap.npy"
SUFFIX_FOR_FLOAT_MMAP = ".float_mmap.npy"
SUFFIX_FOR_READS_MMAP = ".reads_mmap.npy"
SUFFIX_FOR_METADATA = ".metadata.npy"


class MemoryMappedData:
    def __init__(self, int_mmap, float_mmap, num_data: int, reads_mmap, num_reads: int):
        self.int_mmap, self.float_mmap, self.reads_mmap = int_mmap, float_mmap, reads_mmap
        self.num_data, self.num_reads = int(num_data), int(num_reads)
        counts = np.asarray(int_mmap[: self.num_data, : Data.ALT_COUNT.idx + 1]).astype(np.int64)
        ends = np.cumsum(counts[:, Data.REF_COUNT.idx] + counts[:, Data.ALT_COUNT.idx])
        self.read_end_indices = ends.astype(np.uint32)        # memory_mapped_data.py:53-58
        self.read_start_indices = np.concatenate(([0], ends))  # [v] = first row of variant v, [num_data] = rows in use

    def __len__(self) -> int:
        return self.num_data

    def num_bytes(self) -> int:
        return self.int_mmap.nbytes + self.float_mmap.nbytes + (0 if self.reads_mmap is None else self.reads_mmap.nbytes)

    # ---- selection (memory_mapped_data.py:63-101) -----------------------------------------------------------
    def selection_mask(self, num_folds: int = 1, used_folds: Optional[List[int]] = None, label_probs_l=None) -> np.ndarray:
        """Which variants ``generate`` would yield.  With ``label_probs_l`` one ``random.random()`` is drawn per variant of
        the used folds, in order, exactly like the reference's loop (so a seeded run selects the same variants)."""
        keep = np.ones(self.num_data, dtype=bool) if used_folds is None else np.isin(np.arange(self.num_data) % num_folds, list(used_folds))
        if label_probs_l is not None:
            labels = np.asarray(self.int_mmap[: self.num_data, Data.LABEL.idx])
            for idx in np.flatnonzero(keep):
                keep[idx] = random.random() < label_probs_l[int(labels[idx])]
        return keep

    def generate(self, num_folds: int = 1, used_folds: Optional[List[int]] = None, label_probs_l=None) -> Generator[Datum, None, None]:
        """Per-variant view (API parity; the batch loaders never go through it)."""
        keep = self.selection_mask(num_folds, used_folds, label_probs_l)
        for idx in np.flatnonzero(keep):
            r0, r1 = int(self.read_start_indices[idx]), int(self.read_start_indices[idx + 1])
            reads = np.zeros((0, 0), dtype=COMPRESSED_READS_ARRAY_DTYPE) if self.reads_mmap is None else self.reads_mmap[r0:r1]
            yield Datum(self.int_mmap[idx], self.float_mmap[idx], reads, compressed=reads.dtype == COMPRESSED_READS_ARRAY_DTYPE)

    def take(self, keep: np.ndarray) -> "MemoryMappedData":
        """The selected variants, order preserved, as in-memory arrays (reads gathered with one index array)."""
        idx = np.flatnonzero(keep)
        starts, ends = self.read_start_indices[idx], self.read_start_indices[idx + 1]
        lengths = ends - starts
        total = int(lengths.sum())
        # row r of the output comes from starts[v] + (r - first output row of v)
        out_first = np.concatenate(([0], np.cumsum(lengths)))[:-1]
        rows = np.repeat(starts - out_first, lengths) + np.arange(total)
        reads = None if self.reads_mmap is None else np.asarray(self.reads_mmap)[rows]
        return MemoryMappedData(np.asarray(self.int_mmap)[idx], np.asarray(self.float_mmap)[idx], len(idx), reads, total)

    def restrict_to_folds(self, num_folds: int, used_folds: Optional[List[int]] = None, label_probs_l=None) -> "MemoryMappedData":
        if used_folds is None:
            return self
        return self.take(self.selection_mask(num_folds, used_folds, label_probs_l))

    def restrict_to_labeled_only(self) -> "MemoryMappedData":
        from permutect_b200.utils.enums import Label
        labels = np.asarray(self.int_mmap[: self.num_data, Data.LABEL.idx])
        return self.take(labels != Label.UNLABELED)

    # ---- tar file (memory_mapped_data.py:198-285) -----------------------------------------------------------
    def save_to_tarfile(self, output_tarfile: str):
        metadata = np.array([self.num_data, self.int_mmap.shape[-1], self.float_mmap.shape[-1], self.num_reads,
                             0 if self.reads_mmap is None else self.reads_mmap.shape[-1]], dtype=np.uint32)
        with tempfile.TemporaryDirectory() as tmp:
            files = [("metadata" + SUFFIX_FOR_METADATA, None), ("int_array" + SUFFIX_FOR_INT_MMAP, self.int_mmap[: self.num_data]),
                     ("float_array" + SUFFIX_FOR_FLOAT_MMAP, self.float_mmap[: self.num_data])]
            if self.reads_mmap is not None:
                files.append(("reads_array" + SUFFIX_FOR_READS_MMAP, self.reads_mmap[: self.num_reads]))
            with tarfile.open(output_tarfile, "w") as tar:
                for name, array in files:
                    path = os.path.join(tmp, name)
                    if array is None:
                        torch.save(metadata, path)
                    else:
                        np.save(path, np.asarray(array))
                    tar.add(path, arcname=name)

    @classmethod
    def load_from_tarfile(cls, data_tarfile: str) -> "MemoryMappedData":
        temp_dir = tempfile.TemporaryDirectory()
        with tarfile.open(data_tarfile, "r") as tar:
            for member in tar.getmembers():
                if member.isfile():
                    tar.extract(member, path=temp_dir.name)
        files = [os.path.abspath(os.path.join(temp_dir.name, p)) for p in os.listdir(temp_dir.name)]
        pick = lambda suffix: [f for f in files if f.endswith(suffix)]
        meta, ints, floats, reads = pick(SUFFIX_FOR_METADATA), pick(SUFFIX_FOR_INT_MMAP), pick(SUFFIX_FOR_FLOAT_MMAP), pick(SUFFIX_FOR_READS_MMAP)
        if not (len(meta) == 1 and len(ints) == 1 and len(floats) == 1):
            raise ValueError(f"{data_tarfile}: expected one metadata, one int and one float array")
        num_data, int_dim, float_dim, num_reads, reads_dim = (int(x) for x in torch.load(meta[0], weights_only=False)[:5])
        if len(reads) != (0 if num_reads == 0 else 1):
            raise ValueError(f"{data_tarfile}: reads array missing or unexpected")
        # the files may hold more rows than are in use (amortised growth while writing): map the used prefix
        int_mmap = np.load(ints[0], mmap_mode="r")[:num_data]
        float_mmap = np.load(floats[0], mmap_mode="r")[:num_data]
        reads_mmap = None if num_reads == 0 else np.load(reads[0], mmap_mode="r")[:num_reads]
        if int_mmap.shape[1] != int_dim or float_mmap.shape[1] != float_dim or (reads_mmap is not None and reads_mmap.shape[1] != reads_dim):
            raise ValueError(f"{data_tarfile}: array shapes disagree with the metadata")
        result = cls(int_mmap, float_mmap, num_data, reads_mmap, num_reads)
        result._temp_dir = temp_dir      # the maps live in it
        return result
