"""Read-count and logit bins used by the stratified bookkeeping around the model (reference:
permutect/data/count_binning.py).  Ref bins are {0-2}, {3-5}, ...; alt bins {1-3}, {4-6}, ...; logit bins have width 1 on
[-10, 10] so that logit 0 is a bin boundary.  The bin layout is part of the shape of every ``*_slvra`` tensor the
reference's Balancer, Downsampler and metrics hold, so the constants must agree exactly.
"""
from math import floor

import torch
from torch import Tensor

MAX_REF_COUNT, MIN_ALT_COUNT, MAX_ALT_COUNT = 10, 1, 15     # count_binning.py:9-11
MIN_LOGIT, MAX_LOGIT = -10, 10                              # :14-15
COUNT_BIN_SKIP, LOGIT_BIN_SKIP = 3, 1                       # :17, :25
NUM_REF_COUNT_BINS = MAX_REF_COUNT // COUNT_BIN_SKIP + 1                       # 4
NUM_ALT_COUNT_BINS = (MAX_ALT_COUNT - MIN_ALT_COUNT) // COUNT_BIN_SKIP + 1     # 5
NUM_LOGIT_BINS = floor((MAX_LOGIT - MIN_LOGIT) / LOGIT_BIN_SKIP) + 1           # 21
ALT_COUNT_BIN_BOUNDS = [MIN_ALT_COUNT + COUNT_BIN_SKIP * b for b in range(NUM_ALT_COUNT_BINS + 1)]
REF_COUNT_BIN_BOUNDS = [COUNT_BIN_SKIP * b for b in range(NUM_REF_COUNT_BINS + 1)]


# ---- tensors ------------------------------------------------------------------------------------------------
def ref_count_bin_indices(counts: Tensor) -> Tensor:
    return torch.div(counts.clamp(max=MAX_REF_COUNT), COUNT_BIN_SKIP, rounding_mode="floor")


def alt_count_bin_indices(counts: Tensor) -> Tensor:
    return torch.div(counts.clamp(max=MAX_ALT_COUNT) - MIN_ALT_COUNT, COUNT_BIN_SKIP, rounding_mode="floor")


def logit_bin_indices(logits: Tensor) -> Tensor:
    return torch.div(logits.clamp(min=MIN_LOGIT, max=MAX_LOGIT) - MIN_LOGIT, LOGIT_BIN_SKIP, rounding_mode="floor").long()


def logits_from_bin_indices(bins: Tensor) -> Tensor:
    return (MIN_LOGIT + LOGIT_BIN_SKIP / 2) + LOGIT_BIN_SKIP * bins


def counts_from_ref_bin_indices(bins: Tensor) -> Tensor:
    return COUNT_BIN_SKIP * bins + COUNT_BIN_SKIP // 2


def counts_from_alt_bin_indices(bins: Tensor) -> Tensor:
    return MIN_ALT_COUNT + COUNT_BIN_SKIP * bins + COUNT_BIN_SKIP // 2


# ---- scalars ------------------------------------------------------------------------------------------------
def cap_ref_count(count: int) -> int:
    return min(count, MAX_REF_COUNT)


def cap_alt_count(count: int) -> int:
    return min(count, MAX_ALT_COUNT)


def ref_count_bin_index(count: int) -> int:
    return count // COUNT_BIN_SKIP


def alt_count_bin_index(count: int) -> int:
    return (count - MIN_ALT_COUNT) // COUNT_BIN_SKIP


def count_from_ref_bin_index(b: int) -> int:
    return COUNT_BIN_SKIP * b + COUNT_BIN_SKIP // 2


def count_from_alt_bin_index(b: int) -> int:
    return MIN_ALT_COUNT + COUNT_BIN_SKIP * b + COUNT_BIN_SKIP // 2


def round_ref_count_to_bin_center(count: int) -> int:
    return count_from_ref_bin_index(ref_count_bin_index(count))


def round_alt_count_to_bin_center(count: int) -> int:
    return count_from_alt_bin_index(alt_count_bin_index(count))


def top_of_logit_bin(b: int) -> float:
    return MIN_LOGIT + (b + 1) * LOGIT_BIN_SKIP


def ref_count_bin_name(b: int) -> str:
    return str(COUNT_BIN_SKIP * b + (COUNT_BIN_SKIP - 1) // 2)


def alt_count_bin_name(b: int) -> str:
    return str(MIN_ALT_COUNT + COUNT_BIN_SKIP * b + (COUNT_BIN_SKIP - 1) // 2)


def logit_bin_name(b: int) -> str:
    return f"{MIN_LOGIT + (b + 0.5) * LOGIT_BIN_SKIP:.1f}"
