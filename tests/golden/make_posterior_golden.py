"""Golden fixture for the inference caller tail (filter_variants.py:302-320 + MemoryMappedData.from_generator,
memory_mapped_data.py:288-348), generated from the UNMODIFIED reference's Datum class in the build container.

Writes tests/golden/posterior_tail.npz: inputs (int16 / fp16 side arrays, fp32 logits and embeddings) and the two
arrays the reference leaves in the posterior memory map for them.   Usage: python tests/golden/make_posterior_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "oracle", "ref_stubs"), "/root/reference"]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from permutect.data.datum import COMPRESSED_READS_ARRAY_DTYPE, Data, Datum  # noqa: E402


def main():
    rng = np.random.default_rng(123)
    B, L2, I, E = 257, 42, 71, 10
    int_array = rng.integers(-3000, 3000, (B, 16 + L2)).astype(np.int16)
    int_array[:, 0], int_array[:, 1] = rng.integers(0, 11, B), rng.integers(1, 16, B)
    float_array = rng.normal(0, 3, (B, 6 + I)).astype(np.float16)
    float_array[::7, 2] = np.nan                                   # unset AF slots are NaN in real data (datum.py:180-183)
    logits = (rng.normal(0, 9, B)).astype(np.float32)
    logits[:6] = [0.0, -0.0, 20.0, -19.998, 1e-5, 7.00390625]      # fp16 rounding boundaries (quirk Q6)
    emb = rng.normal(0, 2, (B, E)).astype(np.float32)
    logits_t, emb_t = torch.from_numpy(logits), torch.from_numpy(emb)

    # --- verbatim loop body of generate_posterior_data (filter_variants.py:302-320) ---
    data = []
    for ia, fa, logit, embedding in zip(int_array, float_array, logits_t.detach().tolist(), emb_t.cpu()):
        empty_reads = np.zeros((0, 0), dtype=COMPRESSED_READS_ARRAY_DTYPE)
        output_datum = Datum(int_array=ia, float_array=fa, reads_re=empty_reads, compressed=True)
        output_datum.set(Data.REF_COUNT, 0)
        output_datum.set(Data.ALT_COUNT, 0)
        output_datum.set(Data.CACHED_ARTIFACT_LOGIT, logit)
        output_datum.set_info_1d(embedding)
        data.append(output_datum)
    # --- what MemoryMappedData.from_generator stores (memory_mapped_data.py:319-338): dtype of the first datum ---
    int_out = np.zeros((B, data[0].get_int_array().shape[-1]), dtype=data[0].get_int_array().dtype)
    float_out = np.zeros((B, data[0].get_float_array().shape[-1]), dtype=data[0].get_float_array().dtype)
    for i, d in enumerate(data):
        int_out[i] = d.get_int_array()
        float_out[i] = d.get_float_array()
    np.savez_compressed(os.path.join(HERE, "posterior_tail.npz"), int_array=int_array, float_array=float_array, logits=logits,
                        embeddings=emb, int_out=int_out, float_out=float_out)
    print("posterior_tail.npz", int_out.dtype, int_out.shape, float_out.dtype, float_out.shape)


if __name__ == "__main__":
    main()
