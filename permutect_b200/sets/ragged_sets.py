"""Batch of ragged sets as a flattened [N, F] tensor plus per-set lengths
(reference: permutect/sets/ragged_sets.py:16-41).

In this implementation the segmented arithmetic (sums, means, broadcasts) happens inside the fused
CUDA kernels; this class is the value type ``ArtifactModel.calculate_features`` returns, and its
set means are the ones the kernel already produced.
"""
from torch import Tensor


class RaggedSets:
    def __init__(self, flattened_tensor_nf: Tensor, lengths_b: Tensor, means_bf: Tensor = None):
        assert lengths_b.dim() == 1
        self.flattened_tensor_nf = flattened_tensor_nf
        self.lengths_b = lengths_b
        self._means_bf = means_bf

    def batch_size(self) -> int:
        return len(self.lengths_b)

    def means_over_sets(self) -> Tensor:
        """ragged_sets.py:144-155 with the default 1e-4 regulariser weight (computed by the forward kernel)."""
        if self._means_bf is None:
            raise RuntimeError("set means are produced by the fused forward kernel; none were attached")
        return self._means_bf
