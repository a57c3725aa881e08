"""In-memory record layout shared with the reference's datasets (reference: permutect/data/datum.py:23-89).

A variant is an int16 array [16 scalars | 2L haplotype codes], a float16 array [6 scalars | info vector]
and a uint8 array of compressed reads [ref+alt rows][7 packed-bit bytes + quantised floats].
"""
import enum

import numpy as np

INTEGER_DTYPE = np.int16
FLOAT_DTYPE = np.float16
COMPRESSED_READS_ARRAY_DTYPE = np.uint8
RAW_READS_ARRAY_DTYPE = np.float16
NUMBER_OF_BYTES_IN_PACKED_READ = 7   # datum.py:38


class Data(enum.Enum):
    """Column indices (datum.py:51-89).  ``int`` columns live in the int16 array, ``float`` in the fp16 array."""
    REF_COUNT = ("int", 0)
    ALT_COUNT = ("int", 1)
    LABEL = ("int", 2)
    VARIANT_TYPE = ("int", 3)
    SOURCE = ("int", 4)
    ORIGINAL_DEPTH = ("int", 5)
    ORIGINAL_ALT_COUNT = ("int", 6)
    ORIGINAL_NORMAL_DEPTH = ("int", 7)
    ORIGINAL_NORMAL_ALT_COUNT = ("int", 8)
    CONTIG = ("int", 9)
    SEQ_ERROR_LOG_LK = ("float", 0)
    NORMAL_SEQ_ERROR_LOG_LK = ("float", 1)
    ALLELE_FREQUENCY = ("float", 2)
    MAF = ("float", 3)
    NORMAL_MAF = ("float", 4)
    CACHED_ARTIFACT_LOGIT = ("float", 5)

    def __init__(self, kind: str, idx: int):
        self.kind = kind
        self.idx = idx


NUM_SCALAR_INT_ELEMENTS = 16     # datum.py:83
HAPLOTYPES_START_IDX = 16
NUM_SCALAR_FLOAT_ELEMENTS = 6    # datum.py:85
INFO_START_IDX = 6


def num_read_features(row_bytes: int) -> int:
    """datum.py:274-284: 8 bits per packed byte plus one float per remaining byte."""
    return 8 * NUMBER_OF_BYTES_IN_PACKED_READ + (row_bytes - NUMBER_OF_BYTES_IN_PACKED_READ)


class Datum:
    """One variant's arrays (datum.py:92-116).  Reads are ref rows followed by alt rows."""

    def __init__(self, int_array: np.ndarray, float_array: np.ndarray, reads_re: np.ndarray = None,
                 compressed: bool = False):
        assert int_array.ndim == 1 and len(int_array) >= NUM_SCALAR_INT_ELEMENTS
        assert float_array.ndim == 1 and len(float_array) >= NUM_SCALAR_FLOAT_ELEMENTS
        self.int_array = np.asarray(int_array).astype(INTEGER_DTYPE)
        self.float_array = np.asarray(float_array).astype(FLOAT_DTYPE)
        self.reads_re = np.zeros((0, 0), dtype=RAW_READS_ARRAY_DTYPE) if reads_re is None else reads_re
        assert self.reads_re.dtype == (COMPRESSED_READS_ARRAY_DTYPE if compressed else RAW_READS_ARRAY_DTYPE)

    def get(self, field: Data):
        return self.int_array[field.idx] if field.kind == "int" else self.float_array[field.idx]

    def get_int_array(self):
        return self.int_array

    def get_float_array(self):
        return self.float_array

    def get_ref_reads_re(self):
        return self.reads_re[: len(self.reads_re) - int(self.int_array[Data.ALT_COUNT.idx])]

    def get_alt_reads_re(self):
        return self.reads_re[len(self.reads_re) - int(self.int_array[Data.ALT_COUNT.idx]):]

    def get_reads_array_re(self):
        return self.reads_re

    @classmethod
    def from_posterior_record(cls, int_array: np.ndarray, float_array: np.ndarray, empty_reads: np.ndarray) -> "Datum":
        """A posterior record (filter_variants.py:302-320) keeps the fp32 float array np.hstack gave it in the
        reference (datum.py:239-240); the constructor's fp16 cast must not be applied to it."""
        self = cls.__new__(cls)
        self.int_array = np.asarray(int_array, dtype=INTEGER_DTYPE)
        self.float_array = np.asarray(float_array)
        self.reads_re = empty_reads
        return self
