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
