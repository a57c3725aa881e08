"""In-memory record layout shared with the reference's datasets (reference: permutect/data/datum.py:23-89).

A variant is an int16 array [16 scalars | 2L haplotype codes], a float16 array [6 scalars | info vector]
and a uint8 array of compressed reads [ref+alt rows][7 packed-bit bytes + quantised floats].
"""
import enum

import numpy as np

INTEGER_DTYPE = np.int16
LARGE_INTEGER_DTYPE = np.uint32      # stored as TWO consecutive int16 columns (datum.py:24,41-47)
FLOAT_DTYPE = np.float16
BIGGEST_UINT16, BIGGEST_INT16 = 65535, 32767   # datum.py:28-29 (the pair is a base-65535 number in the reference, sic)
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
    POSITION = ("large", 10)
    REF_ALLELE_AS_BASE_5 = ("large", 12)
    ALT_ALLELE_AS_BASE_5 = ("large", 14)
    SEQ_ERROR_LOG_LK = ("float", 0)
    NORMAL_SEQ_ERROR_LOG_LK = ("float", 1)
    ALLELE_FREQUENCY = ("float", 2)
    MAF = ("float", 3)
    NORMAL_MAF = ("float", 4)
    CACHED_ARTIFACT_LOGIT = ("float", 5)

    def __init__(self, kind: str, idx: int):
        self.kind = kind
        self.idx = idx
        # the reference's enum carries the numpy dtype instead of a kind (datum.py:76-78); both spellings work here
        self.dtype = {"int": INTEGER_DTYPE, "large": LARGE_INTEGER_DTYPE, "float": FLOAT_DTYPE}[kind]


def field_kind(field) -> str:
    """'int' / 'large' / 'float' for a column descriptor of this module or of the reference (permutect.data.datum.Data)."""
    kind = getattr(field, "kind", None)
    if kind is not None:
        return kind
    dt = np.dtype(field.dtype)
    return "int" if dt == INTEGER_DTYPE else ("large" if dt == LARGE_INTEGER_DTYPE else "float")


def uint32_from_two_int16s(int16_1, int16_2):
    """datum.py:45-47; works on Python ints, numpy arrays and torch tensors alike."""
    return BIGGEST_UINT16 * (int16_1 + (BIGGEST_INT16 + 1)) + (int16_2 + (BIGGEST_INT16 + 1))


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
        kind = field_kind(field)
        if kind == "int":
            return self.int_array[field.idx]
        if kind == "large":
            return uint32_from_two_int16s(int(self.int_array[field.idx]), int(self.int_array[field.idx + 1]))
        return self.float_array[field.idx]

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
