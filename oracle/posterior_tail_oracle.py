"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's CPU legs, never by the product path).

numpy restatement of the inference caller tail: the per-variant loop of generate_posterior_data
(permutect/tools/filter_variants.py:302-320) followed by what MemoryMappedData.from_generator stores for each Datum
(permutect/data/memory_mapped_data.py:319-338).  Pinned against tests/golden/posterior_tail.npz, which was generated
from the unmodified reference's Datum class (tests/golden/make_posterior_golden.py)."""
import numpy as np

REF_COUNT_IDX, ALT_COUNT_IDX = 0, 1          # datum.py:53-54
CACHED_ARTIFACT_LOGIT_IDX = 5                # datum.py:76
INFO_START_IDX = 6                           # datum.py:89


def posterior_arrays(int_array: np.ndarray, float_array: np.ndarray, logits: np.ndarray, embeddings: np.ndarray):
    """int_array int16 [B, 16+2L], float_array fp16 [B, 6+I], logits fp32 [B], embeddings fp32 [B, E] ->
    (int16 [B, 16+2L], fp32 [B, 6+E]).

    * the counts are zeroed (filter_variants.py:315-316), everything else of the int array is kept;
    * the logit is written into the fp16 float array (datum.py:207-208: rounds to fp16, quirk Q6) BEFORE the info
      block is replaced; np.hstack of that fp16 head with the fp32 embedding row promotes the whole array to fp32
      (datum.py:239-240), and that dtype is what the memory map is created with (memory_mapped_data.py:321)."""
    int_out = np.array(int_array, dtype=np.int16, copy=True)
    int_out[:, REF_COUNT_IDX] = 0
    int_out[:, ALT_COUNT_IDX] = 0
    head = np.array(float_array[:, :INFO_START_IDX], dtype=np.float16, copy=True)
    head[:, CACHED_ARTIFACT_LOGIT_IDX] = np.asarray(logits, dtype=np.float64).astype(np.float16)
    float_out = np.hstack((head.astype(np.float32), np.asarray(embeddings, dtype=np.float32)))
    return int_out, float_out
