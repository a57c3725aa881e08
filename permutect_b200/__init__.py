"""permutect_b200: the ArtifactModel hot path of broadinstitute/permutect on NVIDIA B200.

Python mirrors the reference's model surface (``permutect.architecture.artifact_model``); all
per-read / per-variant arithmetic runs in hand-written sm_100a kernels behind the C-ABI declared in
``include/permutect_b200.h`` (``permutect_b200/csrc``).  There is no CPU or eager-PyTorch fallback:
without the built library and a CUDA device the compute entry points raise.
"""
from permutect_b200 import parameters as _parameters  # noqa: F401  (installs the pickle alias for .pt files)

__version__ = "0.1.0"
