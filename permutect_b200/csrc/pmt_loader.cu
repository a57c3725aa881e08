// Device-side DownsampledBatch (reference: permutect/data/batch.py:383-459).
//
// The reference draws per-read Bernoulli keep masks, forces one alt read per variant, derives the new
// counts with segment_reduce and the gather indices with nonzero() (a host synchronisation).  Here the
// same decisions come from a counter-based hash of (seed, read row), evaluated twice: once to count
// (per variant), once -- after an exclusive scan of the counts -- to write the kept row indices.
// No host round trip, deterministic in the seed.
//
// Quirk Q1 (batch.py:436-439) is preserved: kept ALT entries index the alt block without the ref-block
// offset unless `offset_alt_rows` is set.
#include "pmt_host.h"

namespace pmt {

__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long row) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (row + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);   // 24 random bits -> [0, 1)
}

// alt rows hash with (total_ref + alt row) so ref and alt streams never coincide
__global__ void downsample_count_kernel(const long long* __restrict__ ref_off, const long long* __restrict__ alt_off,
                                        const float* __restrict__ ref_fracs, const float* __restrict__ alt_fracs, int B,
                                        unsigned long long seed, int random_int, long long* __restrict__ new_ref,
                                        long long* __restrict__ new_alt) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= B) return;
  const long long total_ref = ref_off[B];
  const float pr = ref_fracs[v], pa = alt_fracs[v];
  long long nr = 0, na = 0;
  for (long long r = ref_off[v]; r < ref_off[v + 1]; ++r) nr += hash_uniform(seed, (unsigned long long)r) < pr;
  const long long a0 = alt_off[v], a1 = alt_off[v + 1];
  const long long forced = a1 > a0 ? a1 - (random_int % (a1 - a0)) - 1 : -1;   // batch.py:418-425
  for (long long a = a0; a < a1; ++a)
    na += (a == forced) || (hash_uniform(seed, (unsigned long long)(total_ref + a)) < pa);
  new_ref[v] = nr;
  new_alt[v] = na;
}

__global__ void downsample_fill_kernel(const long long* __restrict__ ref_off, const long long* __restrict__ alt_off,
                                       const float* __restrict__ ref_fracs, const float* __restrict__ alt_fracs, int B,
                                       unsigned long long seed, int random_int, const long long* __restrict__ new_ref_off,
                                       const long long* __restrict__ new_alt_off, int offset_alt_rows,
                                       long long* __restrict__ read_indices) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= B) return;
  const long long total_ref = ref_off[B], new_total_ref = new_ref_off[B];
  const float pr = ref_fracs[v], pa = alt_fracs[v];
  long long w = new_ref_off[v];
  for (long long r = ref_off[v]; r < ref_off[v + 1]; ++r)
    if (hash_uniform(seed, (unsigned long long)r) < pr) read_indices[w++] = r;
  const long long a0 = alt_off[v], a1 = alt_off[v + 1];
  const long long forced = a1 > a0 ? a1 - (random_int % (a1 - a0)) - 1 : -1;
  w = new_total_ref + new_alt_off[v];
  for (long long a = a0; a < a1; ++a)
    if ((a == forced) || (hash_uniform(seed, (unsigned long long)(total_ref + a)) < pa))
      read_indices[w++] = offset_alt_rows ? total_ref + a : a;
}

}  // namespace pmt

using namespace pmt;

extern "C" int pmt_downsample_counts(const int64_t* ref_off, const int64_t* alt_off, const float* ref_fracs,
                                     const float* alt_fracs, int32_t n_variants, uint64_t seed, int32_t random_int,
                                     int64_t* new_ref_counts, int64_t* new_alt_counts, void* stream) {
  PMT_CHECK(n_variants > 0, "empty batch");
  downsample_count_kernel<<<(n_variants + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ref_off), reinterpret_cast<const long long*>(alt_off), ref_fracs, alt_fracs,
      n_variants, seed, random_int, reinterpret_cast<long long*>(new_ref_counts), reinterpret_cast<long long*>(new_alt_counts));
  cudaError_t e = cudaGetLastError();
  PMT_CHECK(e == cudaSuccess, "pmt_downsample_counts launch failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int pmt_downsample_fill(const int64_t* ref_off, const int64_t* alt_off, const float* ref_fracs,
                                   const float* alt_fracs, int32_t n_variants, uint64_t seed, int32_t random_int,
                                   const int64_t* new_ref_off, const int64_t* new_alt_off, int32_t offset_alt_rows,
                                   int64_t* read_indices, void* stream) {
  PMT_CHECK(n_variants > 0, "empty batch");
  downsample_fill_kernel<<<(n_variants + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ref_off), reinterpret_cast<const long long*>(alt_off), ref_fracs, alt_fracs,
      n_variants, seed, random_int, reinterpret_cast<const long long*>(new_ref_off),
      reinterpret_cast<const long long*>(new_alt_off), offset_alt_rows, reinterpret_cast<long long*>(read_indices));
  cudaError_t e = cudaGetLastError();
  PMT_CHECK(e == cudaSuccess, "pmt_downsample_fill launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Inference caller tail: the per-variant loop of generate_posterior_data (filter_variants.py:302-320) as one pass.
// int_out = the variant's int16 record with both counts zeroed; float_out (fp32, as np.hstack promotes it,
// datum.py:239-240) = the six fp16 scalars with the fp16-rounded artifact logit in slot 5 (quirk Q6), then the embedding.
// ------------------------------------------------------------------------------------------------
// fp16 -> fp32 keeping a NaN's payload the way numpy's astype does (hardware conversion canonicalises NaNs)
__device__ __forceinline__ float half_bits_to_float(unsigned short h) {
  if ((h & 0x7C00u) == 0x7C00u && (h & 0x03FFu) != 0u)
    return __uint_as_float(((unsigned)(h & 0x8000u) << 16) | 0x7F800000u | ((unsigned)(h & 0x03FFu) << 13));
  return __half2float(__ushort_as_half(h));
}
// fp32 -> fp16 -> fp32 (round to nearest even); a NaN becomes numpy's quiet NaN 0x7e00
__device__ __forceinline__ float round_through_half(float x) {
  if (x != x) return __uint_as_float((__float_as_uint(x) & 0x80000000u) | 0x7FC00000u);
  return __half2float(__float2half_rn(x));
}

__global__ void pack_posterior_kernel(const int16_t* __restrict__ int_in, long long int_stride, int n_int,
                                      const __half* __restrict__ float_in, long long float_stride,
                                      const float* __restrict__ logits, const float* __restrict__ features, int E, int n_variants,
                                      int16_t* __restrict__ int_out, float* __restrict__ float_out) {
  const int wf = 6 + E;
  const long long total = (long long)n_variants * (n_int + wf);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long v = idx / (n_int + wf);
    const int c = (int)(idx - v * (n_int + wf));
    if (c < n_int) {
      int_out[v * n_int + c] = c < 2 ? (int16_t)0 : int_in[v * int_stride + c];
    } else {
      const int f = c - n_int;
      float val;
      if (f == 5) val = round_through_half(logits[v]);
      else if (f < 6) val = half_bits_to_float(__half_as_ushort(float_in[v * float_stride + f]));
      else val = features[v * E + (f - 6)];
      float_out[v * wf + f] = val;
    }
  }
}

extern "C" int pmt_pack_posterior(const int16_t* int_array, int64_t int_stride, int32_t n_int_columns, const void* float_array_f16,
                                  int64_t float_stride, const float* logits_b, const float* features_be, int32_t d_feat,
                                  int32_t n_variants, int16_t* int_out, float* float_out, void* stream) {
  if (n_variants <= 0) return 0;
  const long long total = (long long)n_variants * (n_int_columns + 6 + d_feat);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_posterior_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      int_array, int_stride, n_int_columns, reinterpret_cast<const __half*>(float_array_f16), float_stride, logits_b, features_be,
      d_feat, n_variants, int_out, float_out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
// Batch assembly from the dataset's memory maps (reads_dataset.py:109-196 + Batch.__init__, batch.py:41-62).
// The reads memory map stores, variant after variant, the ref rows then the alt rows of each variant
// (memory_mapped_data.py:36-44); a Batch wants all ref rows of all variants, then all alt rows (batch.py:45-47).
// Instead of re-stacking rows on the host, the contiguous slice of the memory map is shipped as it is and this
// kernel writes the gather indices batch row -> slice row; the read kernels consume them as they do a
// DownsampledBatch's read_indices.
// ------------------------------------------------------------------------------------------------
__global__ void dataset_read_indices_kernel(const long long* __restrict__ ref_off, const long long* __restrict__ alt_off,
                                            int n_variants, long long* __restrict__ read_indices) {
  const long long total_ref = ref_off[n_variants];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_variants; v += gridDim.x * blockDim.x) {
    const long long r0 = ref_off[v], a0 = alt_off[v];
    const long long nr = ref_off[v + 1] - r0, na = alt_off[v + 1] - a0;
    const long long start = r0 + a0;                     // rows of all earlier variants
    for (long long i = 0; i < nr; ++i) read_indices[r0 + i] = start + i;
    for (long long j = 0; j < na; ++j) read_indices[total_ref + a0 + j] = start + nr + j;
  }
}

extern "C" int pmt_dataset_read_indices(const int64_t* ref_off, const int64_t* alt_off, int32_t n_variants, int64_t* read_indices,
                                        void* stream) {
  if (n_variants <= 0) return 0;
  int blocks = (n_variants + 127) / 128;
  if (blocks > 148 * 8) blocks = 148 * 8;
  dataset_read_indices_kernel<<<blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ref_off), reinterpret_cast<const long long*>(alt_off), n_variants,
      reinterpret_cast<long long*>(read_indices));
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
