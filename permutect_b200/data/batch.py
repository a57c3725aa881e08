"""Ragged batch of variants for the fused kernels (reference: permutect/data/batch.py:40-183, 383-459).

Differences from the reference that matter for speed, none for results:
  * reads stay COMPRESSED (uint8 [R, 12], 12 B/read) on the host, over PCIe and in HBM; the decode
    of batch.py:51-56 (unpackbits + wrapped ``(u8-128)/32``) happens inside the read kernel.  The
    reference ships the fp32 expansion (244 B/read).
  * int / float side arrays keep their on-disk dtypes (int16 / fp16, datum.py:23-25).
  * ragged structure is carried as exclusive prefix sums (``ref_off`` / ``alt_off``) computed on the
    device, so no ``.item()`` synchronisation is needed anywhere on the hot path.
Row order is the reference's: all ref reads of all variants, then all alt reads (batch.py:45-47).
"""
from __future__ import annotations

import copy
import enum
from typing import List, Optional

import numpy as np
import torch
from torch import Tensor

from permutect_b200.data import count_binning as bins
from permutect_b200.data.datum import (COMPRESSED_READS_ARRAY_DTYPE, HAPLOTYPES_START_IDX, INFO_START_IDX, Data, Datum,
                                       field_kind, uint32_from_two_int16s)
from permutect_b200.engine import library as L
from permutect_b200.utils.enums import Label, Variation


class Batch:
    def __init__(self, data: List[Datum]):
        """Collate (batch.py:41-62): vstack side arrays; all ref rows, then all alt rows."""
        int_array = np.vstack([d.get_int_array() for d in data])
        float_array = np.vstack([d.get_float_array() for d in data])
        reads = np.vstack([d.get_ref_reads_re() for d in data] + [d.get_alt_reads_re() for d in data])
        self._init_from_arrays(int_array, float_array, reads)

    @classmethod
    def from_arrays(cls, int_array: np.ndarray, float_array: np.ndarray, reads: np.ndarray) -> "Batch":
        """Bulk constructor: ``reads`` already in batch order (all ref rows by variant, then all alt rows)."""
        self = cls.__new__(cls)
        self._init_from_arrays(int_array, float_array, reads)
        return self

    @classmethod
    def from_dataset_slice(cls, int_array: np.ndarray, float_array: np.ndarray, reads: np.ndarray) -> "Batch":
        """Batch cut from a dataset's memory maps: ``reads`` is the CONTIGUOUS slice of the reads map for these variants,
        still in dataset order (variant after variant: its ref rows, then its alt rows; memory_mapped_data.py:36-44).
        No host re-stacking: on the device pmt_dataset_read_indices maps batch rows (batch.py:45-47) onto the slice."""
        self = cls.__new__(cls)
        self._init_from_arrays(int_array, float_array, reads)
        self._dataset_order = True
        return self

    def _init_from_arrays(self, int_array, float_array, reads):
        assert int_array.dtype == np.int16 and float_array.dtype == np.float16
        self.int_tensor = torch.from_numpy(np.ascontiguousarray(int_array))
        self.float_tensor = torch.from_numpy(np.ascontiguousarray(float_array))
        self.reads_are_compressed = reads.dtype == COMPRESSED_READS_ARRAY_DTYPE
        self.reads = torch.from_numpy(np.ascontiguousarray(reads))
        counts = int_array[:, :2].astype(np.int64)
        assert counts[:, 0].sum() + counts[:, 1].sum() == len(reads), "read rows do not match ref+alt counts"
        self.max_rows_per_variant = int((counts[:, 0] + counts[:, 1]).max()) if len(counts) else 0
        self.read_indices = None
        self._dataset_order = False
        self._finish_initialization_from_arrays()

    def _finish_initialization_from_arrays(self):
        self._size = len(self.int_tensor)
        self._offsets = None
        self._decoded = None
        self._device_counts = None      # (ref_counts, alt_counts) int64 when they differ from the int array (downsampling)
        self.lazy_batch_indices = {False: None, True: None}

    # ---- reference accessors (batch.py:68-133,176-182) ----------------------------------------------
    def batch_indices(self, use_original_counts: bool = False) -> "BatchIndices":
        """batch.py:68-85: the (source, label, variant type, ref bin, alt bin) strata of every variant, cached; built from
        the int16 columns on whichever device the batch lives on."""
        cached = self.lazy_batch_indices[use_original_counts]
        if cached is None:
            if use_original_counts:
                ref_counts = self.get(Data.ORIGINAL_DEPTH) - self.get(Data.ORIGINAL_ALT_COUNT)
                alt_counts = self.get(Data.ORIGINAL_ALT_COUNT)
            else:
                ref_counts, alt_counts = self.get(Data.REF_COUNT), self.get(Data.ALT_COUNT)
            cached = BatchIndices(sources=self.get(Data.SOURCE), labels=self.get(Data.LABEL), var_types=self.get(Data.VARIANT_TYPE),
                                  ref_counts=ref_counts, alt_counts=alt_counts)
            self.lazy_batch_indices[use_original_counts] = cached
        return cached

    def get(self, data_field) -> torch.Tensor:
        """batch.py:87-97.  ``data_field`` is a column descriptor of this package or of the reference
        (``permutect.data.datum.Data``): anything with ``idx`` and ``kind`` or ``dtype``.  Integer columns come back int64
        and float columns fp32, the dtypes the reference's batch tensors have."""
        kind, idx = field_kind(data_field), data_field.idx
        if self._device_counts is not None and kind == "int" and idx in (Data.REF_COUNT.idx, Data.ALT_COUNT.idx):
            return self._device_counts[0 if idx == Data.REF_COUNT.idx else 1]
        if kind == "int":
            return self.int_tensor[:, idx].long()
        if kind == "large":
            return uint32_from_two_int16s(self.int_tensor[:, idx].long(), self.int_tensor[:, idx + 1].long())
        return self.float_tensor[:, idx].float()

    def get_training_labels(self) -> torch.Tensor:
        labels = self.int_tensor[:, Data.LABEL.idx]
        return 1.0 * (labels == Label.ARTIFACT) + 0.5 * (labels == Label.UNLABELED)

    def get_is_labeled_mask(self) -> torch.Tensor:
        return (self.int_tensor[:, Data.LABEL.idx] != Label.UNLABELED).int()

    def get_info_be(self) -> torch.Tensor:
        return self.float_tensor[:, INFO_START_IDX:].float()

    def get_haplotypes_bs(self) -> torch.Tensor:
        return self.int_tensor[:, HAPLOTYPES_START_IDX:].long()

    def get_one_hot_haplotypes_bcs(self) -> torch.Tensor:
        """batch.py:115-130 (API parity only; the CNN kernel builds the one-hot image in shared memory)."""
        haps = self.get_haplotypes_bs()
        one_hot = torch.nn.functional.one_hot(haps, num_classes=5).permute(0, 2, 1)
        return one_hot.reshape(len(haps), 10, haps.shape[1] // 2)

    def get_reads_re(self) -> torch.Tensor:
        """Decoded reads [N, F] float32 (batch.py:132-133), produced on demand by the decode kernel."""
        if self._decoded is None:
            self._decoded = self._decode()
        if self.read_indices is None:
            return self._decoded
        idx = self.read_indices
        if self._device_counts is not None:      # downsampled: only the first sum(counts) indices are meaningful
            n_rows = int(self._device_counts[0].sum() + self._device_counts[1].sum())   # API-parity accessor: may synchronise
            idx = idx[:n_rows]
        return self._decoded[idx]

    def _decode(self) -> torch.Tensor:
        if not self.reads_are_compressed:
            return self.reads.float()
        if self.reads.device.type != "cuda":
            raise RuntimeError("read decode runs on the GPU; call batch.copy_to(cuda_device) first")
        lib = L.load()
        n, row_bytes = self.reads.shape
        out = torch.empty((n, 56 + row_bytes - 7), dtype=torch.float32, device=self.reads.device)
        L.check(lib.pmt_decode_reads(self.reads.data_ptr(), n, row_bytes, out.data_ptr(),
                                     torch.cuda.current_stream(self.reads.device).cuda_stream))
        return out

    def get_list_of_reads_re(self):
        """batch.py:137-153: each variant's decoded reads (its ref rows, then its alt rows) as one numpy array."""
        ref_counts, alt_counts = (c.cpu() for c in self.counts())
        reads = self.get_reads_re()
        total_ref = int(ref_counts.sum())
        ref_list = torch.tensor_split(reads[:total_ref], torch.cumsum(ref_counts, dim=0)[:-1])
        alt_list = torch.tensor_split(reads[total_ref:], torch.cumsum(alt_counts, dim=0)[:-1])
        return [torch.vstack((refs, alts)).cpu().numpy() for refs, alts in zip(ref_list, alt_list)]

    def get_int_array_be(self) -> np.ndarray:
        """batch.py:176-177 (int16 here, the records' own dtype; the reference widens to int64 and Datum narrows it back)."""
        result = self.int_tensor.cpu().numpy()
        if self._device_counts is not None:          # DownsampledBatch override, batch.py:451-456
            result = result.copy()
            result[:, Data.REF_COUNT.idx] = self._device_counts[0].cpu().numpy()
            result[:, Data.ALT_COUNT.idx] = self._device_counts[1].cpu().numpy()
        return result

    def get_float_array_be(self) -> np.ndarray:
        """batch.py:179-180."""
        return self.float_tensor.cpu().numpy()

    def size(self) -> int:
        return self._size

    # ---- host <-> device (batch.py:155-174) -------------------------------------------------------
    def pin_memory(self):
        self.int_tensor = self.int_tensor.pin_memory()
        self.float_tensor = self.float_tensor.pin_memory()
        self.reads = self.reads.pin_memory()
        return self

    def copy_to(self, device, dtype=None):
        device = torch.device(device)
        non_blocking = device.type == "cuda"
        new_batch = copy.copy(self)
        new_batch.reads = self.reads.to(device, non_blocking=non_blocking)
        new_batch.int_tensor = self.int_tensor.to(device, non_blocking=non_blocking)
        new_batch.float_tensor = self.float_tensor.to(device, non_blocking=non_blocking)
        if self.read_indices is not None:
            new_batch.read_indices = self.read_indices.to(device, non_blocking=non_blocking)
        new_batch._offsets = None
        new_batch._decoded = None
        new_batch.lazy_batch_indices = {False: None, True: None}     # rebuilt on the new device when first asked for
        new_batch.finalize_on_device()
        return new_batch

    def finalize_on_device(self):
        """Device-side part of batch assembly: gather indices for reads that arrived in dataset order."""
        if not getattr(self, "_dataset_order", False) or self.read_indices is not None or self.reads.device.type != "cuda":
            return
        ref_off, alt_off = self.offsets()
        dev = self.reads.device
        self.read_indices = torch.empty(self.reads.shape[0], dtype=torch.int64, device=dev)
        L.check(L.load().pmt_dataset_read_indices(ref_off.data_ptr(), alt_off.data_ptr(), self._size, self.read_indices.data_ptr(),
                                                  torch.cuda.current_stream(dev).cuda_stream))

    def detach_from_ring(self) -> "Batch":
        """Give this batch its own device memory.  Batches yielded by prefetch_generator are views into a ring of persistent
        buffers that is refilled a few iterations later; call this before keeping one (or a DownsampledBatch of it) longer."""
        for name in ("reads", "int_tensor", "float_tensor", "read_indices"):
            t = getattr(self, name, None)
            if t is not None:
                setattr(self, name, t.clone())
        self._offsets = None
        self._decoded = None
        self.lazy_batch_indices = {False: None, True: None}
        return self

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.reads, self.int_tensor, self.float_tensor))

    # ---- kernel-facing view -----------------------------------------------------------------------
    def counts(self):
        if self._device_counts is not None:
            return self._device_counts
        return self.int_tensor[:, Data.REF_COUNT.idx].long(), self.int_tensor[:, Data.ALT_COUNT.idx].long()

    def offsets(self):
        """Exclusive prefix sums [B+1] (int64) of the ref and alt counts, on the batch's device."""
        if self._offsets is None:
            ref_counts, alt_counts = self.counts()
            off = torch.zeros((2, self._size + 1), dtype=torch.int64, device=ref_counts.device)
            # two 1-D scans (single-pass device scan); a [2, B] scan along dim 1 takes torch's generic kernel, 50x slower
            torch.cumsum(ref_counts.to(torch.int64), dim=0, out=off[0, 1:])
            torch.cumsum(alt_counts.to(torch.int64), dim=0, out=off[1, 1:])
            self._offsets = off
        return self._offsets[0], self._offsets[1]

    def pmt_batch(self) -> L.PmtBatch:
        """Fill the C-ABI PmtBatch (include/permutect_b200.h).  Keeps no host copies of device data."""
        if self.reads.device.type != "cuda":
            raise RuntimeError("the ArtifactModel kernels need the batch on a CUDA device (Batch.copy_to)")
        ref_off, alt_off = self.offsets()
        b = L.PmtBatch()
        b.n_variants = self._size
        if self.reads_are_compressed:
            b.reads_kind = L.READS_U8
        else:
            b.reads_kind = L.READS_F16 if self.reads.dtype == torch.float16 else L.READS_F32
        b.info_kind, b.hap_kind = L.F16, L.I16
        # exact totals come from ref_off[B] / alt_off[B] on the device; the host only knows an upper bound (grid sizing)
        b.n_rows = int(self.read_indices.shape[0] if self.read_indices is not None else self.reads.shape[0])
        b.total_ref = -1
        b.max_rows_per_variant = self.max_rows_per_variant
        b.reads = self.reads.data_ptr()
        b.read_indices = self.read_indices.data_ptr() if self.read_indices is not None else None
        b.ref_off, b.alt_off = ref_off.data_ptr(), alt_off.data_ptr()
        b.info = self.float_tensor.data_ptr() + INFO_START_IDX * self.float_tensor.element_size()
        b.info_stride = self.float_tensor.shape[1]
        b.haplotypes = self.int_tensor.data_ptr() + HAPLOTYPES_START_IDX * self.int_tensor.element_size()
        b.hap_stride = self.int_tensor.shape[1]
        return b


class BatchProperty(enum.IntEnum):
    """Axes of a batch-indexed tensor and the display names of their bins (batch.py:185-201)."""
    SOURCE = (0, None)
    LABEL = (1, [label.name for label in Label])
    VARIANT_TYPE = (2, [var_type.name for var_type in Variation])
    REF_COUNT_BIN = (3, [bins.ref_count_bin_name(i) for i in range(bins.NUM_REF_COUNT_BINS)])
    ALT_COUNT_BIN = (4, [bins.alt_count_bin_name(i) for i in range(bins.NUM_ALT_COUNT_BINS)])
    LOGIT_BIN = (5, [bins.logit_bin_name(i) for i in range(bins.NUM_LOGIT_BINS)])

    def __new__(cls, value, names_list):
        member = int.__new__(cls, value)
        member._value_ = value
        member.names_list = names_list
        return member

    def get_name(self, n: int) -> str:
        return str(n) if self.names_list is None else self.names_list[n]


_LABEL_STRIDE = len(Variation) * bins.NUM_REF_COUNT_BINS * bins.NUM_ALT_COUNT_BINS      # 100
_SOURCE_STRIDE = len(Label) * _LABEL_STRIDE                                              # 300


class BatchIndices:
    """Row-major position of every variant in a ``[source, label, variant type, ref bin, alt bin]`` table
    (batch.py:204-291).  Source is the leading axis, so the position does not depend on the number of sources."""

    def __init__(self, sources: Tensor, labels: Tensor, var_types: Tensor, ref_counts: Tensor, alt_counts: Tensor):
        self.sources, self.labels, self.var_types = sources, labels, var_types
        self.ref_count_bins = bins.ref_count_bin_indices(ref_counts)
        self.alt_count_bins = bins.alt_count_bin_indices(alt_counts)
        self.flattened_idx = (sources * _SOURCE_STRIDE + labels * _LABEL_STRIDE
                              + (var_types * bins.NUM_REF_COUNT_BINS + self.ref_count_bins) * bins.NUM_ALT_COUNT_BINS
                              + self.alt_count_bins)

    def _flattened_idx(self, source_override: Tensor = None, pseudolabels: Tensor = None, logits: Tensor = None) -> Tensor:
        """Positions with the source and / or label replaced (a shift by whole strides), optionally extended by the logit
        bin as the innermost axis (batch.py:230-262)."""
        idx = self.flattened_idx
        if source_override is not None:
            assert len(source_override) == len(self.sources)
            idx = idx + _SOURCE_STRIDE * (source_override - self.sources)
        if pseudolabels is not None:
            assert len(pseudolabels) == len(self.labels)
            idx = idx + _LABEL_STRIDE * (pseudolabels - self.labels)
        if logits is not None:
            idx = bins.logit_bin_indices(logits) + bins.NUM_LOGIT_BINS * idx
        return idx

    def index_into_tensor(self, tens, sources: Tensor = None, labels: Tensor = None, logits: Tensor = None) -> Tensor:
        """result[i] = tens[source[i], label[i], variant type[i], ref bin[i], alt bin[i] (, logit bin[i])] (batch.py:264-278)."""
        assert (logits is None) == (not tens.has_logits()), "Logits used iff batch-indexed tensor has logit dimension."
        return tens.view(-1)[self._flattened_idx(source_override=sources, pseudolabels=labels, logits=logits)]

    def increment_tensor(self, tens, values: Tensor, sources: Tensor = None, labels: Tensor = None, logits: Tensor = None):
        """tens[strata of variant i] += values[i], in place (batch.py:280-291)."""
        assert (logits is None) == (not tens.has_logits()), "Logits used iff batch-indexed tensor has logit dimension."
        idx = self._flattened_idx(source_override=sources, pseudolabels=labels, logits=logits)
        return tens.view(-1).index_add_(dim=0, index=idx, source=values)


class BatchIndexedTensor(Tensor):
    """Sums stratified by source, label, variant type, ref bin, alt bin and optionally logit bin (batch.py:294-380):
    a plain tensor of 5 or 6 dimensions that knows its axes."""

    @staticmethod
    def __new__(cls, data: Tensor):
        return torch.Tensor._make_subclass(cls, data)

    def __init__(self, data: Tensor):
        assert data.dim() in (5, 6), "batch-indexed tensors have either 5 or 6 dimensions"

    def has_logits(self) -> bool:
        return self.dim() == 6

    def num_sources(self) -> int:
        return self.shape[0]

    @classmethod
    def shape_without_logits(cls, num_sources: int):
        return (num_sources, len(Label), len(Variation), bins.NUM_REF_COUNT_BINS, bins.NUM_ALT_COUNT_BINS)

    @classmethod
    def _filled(cls, value: float, num_sources: int, include_logits: bool, device):
        shape = cls.shape_without_logits(num_sources) + ((bins.NUM_LOGIT_BINS,) if include_logits else ())
        return cls(torch.full(shape, value, dtype=torch.float32, device=device))

    @classmethod
    def zeros(cls, num_sources: int, include_logits: bool = False, device=None):
        return cls._filled(0.0, num_sources, include_logits, device)

    @classmethod
    def ones(cls, num_sources: int, include_logits: bool = False, device=None):
        return cls._filled(1.0, num_sources, include_logits, device)

    def resize_sources(self, new_num_sources: int):
        old = self.num_sources()
        shape = self.shape_without_logits(new_num_sources) + ((bins.NUM_LOGIT_BINS,) if self.has_logits() else ())
        self.resize_(shape)
        self[old:] = 0

    def record_datum(self, datum: Datum, value: float = 1.0, grow_source_if_necessary: bool = True):
        assert not self.has_logits(), "this only works when not including logits"
        source = int(datum.get(Data.SOURCE))
        if source >= self.num_sources():
            if not grow_source_if_necessary:
                raise Exception("Datum source doesn't fit.")
            self.resize_sources(source + 1)
        self[source, int(datum.get(Data.LABEL)), int(datum.get(Data.VARIANT_TYPE)),
             bins.ref_count_bin_index(int(datum.get(Data.REF_COUNT))), bins.alt_count_bin_index(int(datum.get(Data.ALT_COUNT)))] += value

    def record(self, batch: Batch, values: Tensor, logits: Tensor = None, use_original_counts: bool = False):
        batch.batch_indices(use_original_counts).increment_tensor(self, values=values, logits=logits)

    def get_marginal(self, *properties: BatchProperty) -> Tensor:
        """Sum over every axis that is not listed (batch.py:371-380)."""
        keep = set(int(p) for p in properties)
        return torch.sum(self, dim=tuple(n for n in range(self.dim()) if n not in keep))


class DownsampledBatch(Batch):
    """Read-downsampled view of a parent batch without copying reads (batch.py:383-459).

    ``read_indices`` lists the kept rows: kept ref rows followed by kept alt rows.  As in the reference
    the alt entries index the alt block WITHOUT the ref-block offset (quirk Q1, batch.py:436-439): a downsampled alt set
    is gathered from the rows of the REF block.  That is the reference's behaviour, so it is the default here (a model
    trained here sees what a model trained there sees); the first default draw of a process says so once.
    ``offset_alt_rows=True`` -- per call, or for every call through ``DownsampledBatch.OFFSET_ALT_ROWS = True`` -- gathers
    the alt rows the keep mask selected (what upstream presumably intended); ``tests/test_downsample_gpu.py`` pins both.
    """

    OFFSET_ALT_ROWS: Optional[bool] = None     # None: the reference's behaviour (quirk Q1) with a one-time warning
    _warned_q1 = False

    def __init__(self, original_batch: Batch, ref_fracs_b: Optional[torch.Tensor] = None,
                 alt_fracs_b: Optional[torch.Tensor] = None, *, read_indices: Optional[torch.Tensor] = None,
                 ref_counts: Optional[torch.Tensor] = None, alt_counts: Optional[torch.Tensor] = None,
                 offset_alt_rows: Optional[bool] = None, seed: Optional[int] = None):
        if offset_alt_rows is None:
            offset_alt_rows = DownsampledBatch.OFFSET_ALT_ROWS
        if offset_alt_rows is None:
            offset_alt_rows = False
            if read_indices is None and not DownsampledBatch._warned_q1:
                DownsampledBatch._warned_q1 = True
                import warnings
                warnings.warn("DownsampledBatch reproduces the reference's alt-row indexing (batch.py:436-439: kept alt indices lack "
                              "the ref-block offset, so downsampled alt sets are gathered from ref rows). Set "
                              "DownsampledBatch.OFFSET_ALT_ROWS = True (or pass offset_alt_rows=True) for the corrected gather, "
                              "False to keep the reference's behaviour without this message.", stacklevel=2)
        self.int_tensor = original_batch.int_tensor
        self.float_tensor = original_batch.float_tensor
        self.reads = original_batch.reads
        self.reads_are_compressed = original_batch.reads_are_compressed
        self.max_rows_per_variant = original_batch.max_rows_per_variant
        self.device = self.int_tensor.device
        self._finish_initialization_from_arrays()
        self._decoded = original_batch._decoded
        if read_indices is not None:       # pre-drawn masks (tests, reproducible benchmarks)
            self.read_indices = read_indices.to(self.device, torch.int64)
            self._device_counts = (ref_counts.to(self.device, torch.int64), alt_counts.to(self.device, torch.int64))
        else:
            from permutect_b200.data.downsample import downsample_on_device
            self.read_indices, ref_c, alt_c, new_off = downsample_on_device(original_batch, ref_fracs_b, alt_fracs_b,
                                                                            offset_alt_rows=offset_alt_rows, seed=seed)
            self._device_counts = (ref_c, alt_c)
            self._offsets = new_off
        self.ref_counts, self.alt_counts = self._device_counts
        self._dataset_order = False
        if original_batch.read_indices is not None:
            # the parent's rows are themselves reached through gather indices (a dataset-order batch): compose them.
            # (entries past the kept rows are zero, a valid index)
            self.read_indices = original_batch.read_indices[self.read_indices]
