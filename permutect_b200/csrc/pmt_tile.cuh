// Read-path tile machinery shared by the forward, long-set and backward kernels.
//
// A tile is up to TILE rows (reads): the ref reads of a run of consecutive variants, padded to a
// multiple of 4 rows, followed by their alt reads.  Each phase below is one step of
// ArtifactModel.calculate_features / FeatureClustering (artifact_model.py:239-297) over the tile.
#pragma once
#include "pmt_device.cuh"

namespace pmt {

struct HeadConst {   // per-CTA constants of the clustering head (feature_clustering.py:82-119)
  float sigma[PMT_MAX_FEAT];
  float c_non, c_out;
  float c_orth[PMT_MAX_CLUSTERS], two_tau2[PMT_MAX_CLUSTERS];
  float log_half_lambda[PMT_MAX_CLUSTERS], shift[PMT_MAX_CLUSTERS], sqrt2_sigma[PMT_MAX_CLUSTERS],
      half_lambda[PMT_MAX_CLUSTERS], two_mu_plus[PMT_MAX_CLUSTERS], logw[PMT_MAX_CLUSTERS];
};

struct TileMeta {
  int nv;             // variants in the tile
  int v0;             // first variant (global index)
  int ref_pad;        // rows [0, ref_pad) are ref rows (incl. padding), rows [ref_pad, rows) alt rows
  int rows;
  int alt_head;       // 1 if the first alt row of the tile's variants is inside this tile (0 for later chunks of a long set)
  int rowvar[TILE];         // local variant of each row, -1 for padding
  long long rowidx[TILE];   // batch row index (position in [0, n_rows)), -1 for padding
  int ref_start[TILE], ref_cnt[TILE], alt_start[TILE], alt_cnt[TILE];  // per local variant: rows inside THIS tile
  float ref_total[TILE], alt_total[TILE];                              // per local variant: set sizes (== cnt unless chunked)
};

struct TileCtx {
  float* X;      // residual stream / read embedding     [PMT_MAX_DIM][LD]
  float* T1;
  float* T2;
  float* sums;   // [TILE][2][sum_w]   per-variant sums / mean fields (ref, alt)
  float* llsum;  // [TILE][2][16]      per-variant log-likelihood sums (alt side used)
  HeadConst* HC;
  TileMeta* M;
  const float* W;      // flat materialised weights
  int rows_used;       // rows rounded up to a multiple of 4
};

struct ReadKernelArgs {
  const float* wflat;
  const float* image;
  PmtBatch batch;
  PmtOutputs out;
  int* claim_counter;
  float* scratch;       // long-set / backward activation scratch (per CTA regions)
  long long scratch_stride;  // floats per CTA
};

// exponentially_modified_gaussian.py:30-55
__device__ __forceinline__ float logerfc(float z) {
  if (z > 5.f) {
    const float z2 = z * z, z4 = z2 * z2, z6 = z2 * z4;
    return -z2 - logf(z * 1.7724538509055160273f) + log1pf(-1.f / (2.f * z2) + 3.f / (4.f * z4) - 15.f / (8.f * z6));
  }
  return logf(fmaxf(erfcf(z), 1.0e-12f));
}
// d logerfc / dz, branch-wise as autograd differentiates the reference's torch.where
__device__ __forceinline__ float dlogerfc(float z) {
  if (z > 5.f) {
    const float z2 = z * z, z3 = z2 * z, z5 = z3 * z2, z7 = z5 * z2;
    const float s = -1.f / (2.f * z2) + 3.f / (4.f * z2 * z2) - 15.f / (8.f * z2 * z2 * z2);
    const float ds = 1.f / z3 - 3.f / z5 + 45.f / (4.f * z7);
    return -2.f * z - 1.f / z + ds / (1.f + s);
  }
  const float e = erfcf(z);
  return e > 1.0e-12f ? -1.1283791670955125739f * expf(-z * z) / e : 0.f;
}

__device__ __forceinline__ float logsumexp2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}

__device__ __forceinline__ void head_constants(const PmtModelDesc& D, const float* W, HeadConst* HC) {
  const int E = D.d_feat, K = D.n_clusters;
  float sum_log_sigma = 0.f, sum_log_2sigma = 0.f;
  for (int e = 0; e < E; ++e) {
    const float s = W[D.sigma_e + e];
    HC->sigma[e] = s; sum_log_sigma += logf(s); sum_log_2sigma += logf(2.f * s);
  }
  HC->c_non = -(E * 0.5f) * LOG_2PI - sum_log_sigma;
  HC->c_out = -(E * 0.5f) * LOG_2PI - sum_log_2sigma;
  for (int k = 0; k < K; ++k) {
    const float tau = W[D.tau_k + k], lam = W[D.lambda_k + k], sg = W[D.emg_sigma_k + k], mu = W[D.mu_k + k];
    HC->c_orth[k] = -((E - 1) * 0.5f) * LOG_2PI - (E - 1) * logf(tau);
    HC->two_tau2[k] = 2.f * tau * tau;
    HC->log_half_lambda[k] = logf(lam / 2.f);
    HC->shift[k] = mu + lam * sg * sg;
    HC->sqrt2_sigma[k] = 1.41421356237309504880f * sg;
    HC->half_lambda[k] = lam / 2.f;
    HC->two_mu_plus[k] = 2.f * mu + lam * sg * sg;
    HC->logw[k] = W[D.logw_k + k];
  }
}

// Per-variant sums of feature rows [f0, f0+nf) of `buf` over the ref rows and the alt rows of each
// variant in the tile (ragged_sets.py:157-158).  out is [nv][2][sw]; accumulate adds to it.
__device__ __forceinline__ void segment_sums(const TileMeta& M, const float* buf, int f0, int nf, float* out, int sw,
                                             bool alt_only, bool accumulate) {
  const int sides = alt_only ? 1 : 2;
  for (int idx = threadIdx.x; idx < M.nv * sides * nf; idx += NTHREADS) {
    const int j = idx / (sides * nf), rem = idx % (sides * nf);
    const int s = alt_only ? 1 : rem / nf, f = rem % nf;
    const int start = s ? M.alt_start[j] : M.ref_start[j], cnt = s ? M.alt_cnt[j] : M.ref_cnt[j];
    const float* p = buf + (f0 + f) * LD + start;
    float sum = 0.f;
    for (int i = 0; i < cnt; ++i) sum += p[i];
    float* o = out + (j * 2 + s) * sw + f;
    *o = accumulate ? *o + sum : sum;
  }
}

// Greedy tile construction: take whole variants from v_cur while ref rows (padded to 4) + alt rows fit.
// Returns the number of variants taken (CTA-uniform); 0 means variant v_cur alone does not fit.
__device__ __forceinline__ int build_tile(const PmtBatch& batch, int v_cur, int v_end, long long total_ref, TileMeta& M) {
  const int tid = threadIdx.x;
  const long long r_base = __ldg(batch.ref_off + v_cur), a_base = __ldg(batch.alt_off + v_cur);
  int fits = 0;
  if (tid < TILE && v_cur + tid + 1 <= v_end) {
    const long long nr = __ldg(batch.ref_off + v_cur + tid + 1) - r_base;
    const long long na = __ldg(batch.alt_off + v_cur + tid + 1) - a_base;
    fits = (((nr + 3) & ~3LL) + na <= TILE) ? 1 : 0;
  }
  const int nv = __syncthreads_count(fits);
  if (nv == 0) return 0;
  if (tid < TILE) { M.rowvar[tid] = -1; M.rowidx[tid] = -1; }
  __syncthreads();
  const long long nr_tot = __ldg(batch.ref_off + v_cur + nv) - r_base;
  const long long na_tot = __ldg(batch.alt_off + v_cur + nv) - a_base;
  const int ref_pad = (int)((nr_tot + 3) & ~3LL);
  if (tid < nv) {
    const long long r0 = __ldg(batch.ref_off + v_cur + tid), r1 = __ldg(batch.ref_off + v_cur + tid + 1);
    const long long a0 = __ldg(batch.alt_off + v_cur + tid), a1 = __ldg(batch.alt_off + v_cur + tid + 1);
    const int rs = (int)(r0 - r_base), rc = (int)(r1 - r0), as = ref_pad + (int)(a0 - a_base), ac = (int)(a1 - a0);
    M.ref_start[tid] = rs; M.ref_cnt[tid] = rc; M.alt_start[tid] = as; M.alt_cnt[tid] = ac;
    M.ref_total[tid] = (float)rc; M.alt_total[tid] = (float)ac;
    for (int i = 0; i < rc; ++i) { M.rowvar[rs + i] = tid; M.rowidx[rs + i] = r0 + i; }
    for (int i = 0; i < ac; ++i) { M.rowvar[as + i] = tid; M.rowidx[as + i] = total_ref + a0 + i; }
  }
  if (tid == 0) { M.nv = nv; M.v0 = v_cur; M.ref_pad = ref_pad; M.rows = ref_pad + (int)na_tot; M.alt_head = 1; }
  __syncthreads();
  return nv;
}

// Chunk c of a single long variant v: virtual rows [c*TILE, (c+1)*TILE) of [ref rows | pad to 4 | alt rows].
__device__ __forceinline__ void build_chunk(const PmtBatch& batch, int v, int c, long long total_ref, TileMeta& M) {
  const int tid = threadIdx.x;
  const long long r0 = __ldg(batch.ref_off + v), r1 = __ldg(batch.ref_off + v + 1);
  const long long a0 = __ldg(batch.alt_off + v), a1 = __ldg(batch.alt_off + v + 1);
  const long long nref = r1 - r0, nalt = a1 - a0, ref_pad = (nref + 3) & ~3LL, total = ref_pad + nalt;
  const long long lo = (long long)c * TILE;
  __syncthreads();
  if (tid < TILE) {
    const long long vr = lo + tid;
    long long idx = -1;
    if (vr < nref) idx = r0 + vr;
    else if (vr >= ref_pad && vr < total) idx = total_ref + a0 + (vr - ref_pad);
    M.rowidx[tid] = idx;
    M.rowvar[tid] = idx >= 0 ? 0 : -1;
  }
  if (tid == 0) {
    const long long hi = lo + TILE;
    const long long ref_lo = lo < nref ? lo : nref, ref_hi = hi < nref ? hi : nref;
    const long long alt_lo = (lo > ref_pad ? lo : ref_pad), alt_hi = (hi < total ? hi : total);
    M.nv = 1; M.v0 = v;
    M.alt_head = (ref_pad >= lo && ref_pad < hi) ? 1 : 0;
    M.ref_pad = (int)(ref_pad <= lo ? 0 : (ref_pad >= hi ? TILE : ref_pad - lo));
    M.rows = (int)((hi < total ? hi : total) - lo);
    M.ref_start[0] = (int)(ref_lo - lo); M.ref_cnt[0] = (int)(ref_hi - ref_lo);
    M.alt_start[0] = alt_hi > alt_lo ? (int)(alt_lo - lo) : 0; M.alt_cnt[0] = alt_hi > alt_lo ? (int)(alt_hi - alt_lo) : 0;
    M.ref_total[0] = (float)nref; M.alt_total[0] = (float)nalt;
  }
  __syncthreads();
}

// batch.py:51-56 + plain_text_data.py:510-511: decode the tile's reads into T1 (feature-major)
__device__ __forceinline__ void tile_decode(const PmtModelDesc& D, const PmtBatch& batch, const TileMeta& M, float* T1) {
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  const int F = D.n_read_features;
  const long long my_idx = M.rowidx[row];
  long long src = -1;
  if (my_idx >= 0) src = batch.read_indices ? __ldg(batch.read_indices + my_idx) : my_idx;
  if (batch.reads_kind == PMT_READS_U8) {
    const int rb = D.read_row_bytes;
    const uint8_t* rp = reinterpret_cast<const uint8_t*>(batch.reads) + src * rb;
    // the packed bytes 0..6 are spread over the parts of a row; the last part also takes the quantised floats
    const int b_lo = (8 * part) / NPART, b_hi = min(7, (8 * (part + 1)) / NPART);
    for (int b = b_lo; b < b_hi; ++b) {
      const unsigned byte = src >= 0 ? __ldg(rp + b) : 0u;
#pragma unroll
      for (int bit = 0; bit < 8; ++bit) T1[(b * 8 + bit) * LD + row] = (float)((byte >> (7 - bit)) & 1u);
    }
    if (part == NPART - 1) {
      for (int b = 7; b < rb; ++b) {
        const unsigned byte = src >= 0 ? __ldg(rp + b) : 128u;
        T1[(56 + b - 7) * LD + row] = (float)((byte + 128u) & 255u) * 0.03125f;   // wraps like the uint8 arithmetic, quirk Q2
      }
    }
  } else {
    for (int f = part; f < F; f += NPART) {
      float v = 0.f;
      if (src >= 0) {
        v = batch.reads_kind == PMT_READS_F16 ? __half2float(reinterpret_cast<const __half*>(batch.reads)[src * F + f])
                                              : reinterpret_cast<const float*>(batch.reads)[src * F + f];
      }
      T1[f * LD + row] = v;
    }
  }
}

// artifact_model.py:243-251: read embedding into X[0..d_read), info/seq embedding of the row's variant into
// X[d_read..d_model).  Leaves X complete after a trailing __syncthreads().
__device__ __forceinline__ void tile_embed(const Plan& P, TileCtx& C, Stage& stage, const PmtBatch& batch,
                                           const float* info_seq, float* scr) {
  const PmtModelDesc& D = P.d;
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  stage.prefetch(P.read_g0);
  tile_decode(D, batch, *C.M, C.T1);
  float* emb = run_mlp(P, D.read_ops, D.n_read_ops, P.read_g0, C.T1, C.X, C.T1, C.T2, stage, C.W, C.rows_used,
                       P.blk_g0, scr, P.scr_read);
  __syncthreads();
  if (emb != C.X) copy_features(emb, C.X, D.d_read);
  const int w = D.d_info + D.d_seq;
  const int my_var = C.M->rowvar[row];
  const float* src = my_var >= 0 ? info_seq + (long long)(C.M->v0 + my_var) * w : nullptr;
  for (int j = part; j < w; j += NPART) C.X[(D.d_read + j) * LD + row] = src ? __ldg(src + j) : 0.f;
  __syncthreads();
}

// LayerNorm statistics of `nf` features of one row held feature-major in buf
__device__ __forceinline__ void row_stats(const float* buf, int nf, int row, float& mean, float& rstd) {
  float m = 0.f;
  for (int f = 0; f < nf; ++f) m += buf[f * LD + row];
  m /= nf;
  float var = 0.f;
  for (int f = 0; f < nf; ++f) { const float d = buf[f * LD + row] - m; var = fmaf(d, d, var); }
  mean = m;
  rstd = rsqrtf(var / nf + LN_EPS);
}

// gated_mlp.py:230-233: LayerNorm of the gating half z2 = T2[H..2H), in place.  Ends with a __syncthreads().
__device__ __forceinline__ void sgu_layernorm(const Plan& P, TileCtx& C, int blk) {
  const PmtBlockOffsets& BO = P.d.blocks[blk];
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  const int H = P.d.d_ffn / 2;
  const float* W = C.W;
  float mean, rstd;
  row_stats(C.T2 + H * LD, H, row, mean, rstd);
  __syncthreads();  // every part (and a preceding save) has read the raw z2 of this row
  const int f_lo = part_lo(H, part), f_hi = part_lo(H, part + 1);
  for (int f = f_lo; f < f_hi; ++f)
    C.T2[(H + f) * LD + row] = (C.T2[(H + f) * LD + row] - mean) * rstd * __ldg(W + BO.ln2_w + f) + __ldg(W + BO.ln2_b + f);
  __syncthreads();
}

// gated_mlp.py:185-190, 230-233: T1 = LN(X); T2[0..d_ffn) = SELU(proj1_s T1); T2[H..2H) <- LN2 in place.
// `z_save` (optional): global image receiving z = T2[0..d_ffn) BEFORE the SGU LayerNorm.
__device__ __forceinline__ void block_phase_a(const Plan& P, TileCtx& C, Stage& stage, int blk, float* z_save) {
  const PmtModelDesc& D = P.d;
  const PmtBlockOffsets& BO = D.blocks[blk];
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  const int Dm = D.d_model;
  const int g1 = P.blk_g0 + 2 * blk;
  const float* W = C.W;
  {
    float mean, rstd;
    row_stats(C.X, Dm, row, mean, rstd);
    const int f_lo = part_lo(Dm, part), f_hi = part_lo(Dm, part + 1);
    for (int f = f_lo; f < f_hi; ++f)
      C.T1[f * LD + row] = (C.X[f * LD + row] - mean) * rstd * __ldg(W + BO.ln_w + f) + __ldg(W + BO.ln_b + f);
  }
  const float* img1 = stage.acquire(g1);
  stage.prefetch(g1 + 1);
  gemm_tile(C.T1, P.gemm[g1], img1, W, C.M->ref_pad, C.T2, EPI_SELU, 1.f, C.rows_used);
  __syncthreads();
  if (z_save) save_rows(C.T2, D.d_ffn, z_save);
  sgu_layernorm(P, C, blk);
}

// gated_mlp.py:236-239, ragged_sets.py:144-155: turn per-variant sums of z2 into the ref / alt mean fields, in place
__device__ __forceinline__ void block_means(const Plan& P, TileCtx& C, int blk) {
  const PmtBlockOffsets& BO = P.d.blocks[blk];
  const int H = P.d.d_ffn / 2;
  const float regw = __ldg(C.W + BO.reg_weight) + 0.25f;
  for (int idx = threadIdx.x; idx < C.M->nv * 2 * H; idx += NTHREADS) {
    const int j = idx / (2 * H), s = (idx / H) & 1, f = idx % H;
    float* p = C.sums + (j * 2 + s) * P.sum_w + f;
    if (s == 0) *p = (*p + regw * __ldg(C.W + BO.regularizer + f)) / (C.M->ref_total[j] + regw);
    else *p = *p / (C.M->alt_total[j] + 1e-4f);
  }
  __syncthreads();
}

// gated_mlp.py:243-251: gate value of feature f for `row` (z2n = normalised z2 of that row/feature)
__device__ __forceinline__ float gate_value(const Plan& P, const TileCtx& C, const PmtBlockOffsets& BO, int row, int f,
                                            float z2n, int my_var, bool is_alt, float alpha, float beta, float gamma) {
  float gate = z2n * alpha + 1.f;
  if (my_var >= 0) {
    const float m_ref = C.sums[(my_var * 2 + 0) * P.sum_w + f];
    if (is_alt) gate = gate + beta * C.sums[(my_var * 2 + 1) * P.sum_w + f] + gamma * m_ref;
    else gate = gate + beta * m_ref;
  }
  return gate;
}

// gated_mlp.py:196-200, 243-251: T1[0..H) = z1 * gate; X += proj2_s T1
__device__ __forceinline__ void block_phase_b(const Plan& P, TileCtx& C, Stage& stage, int blk, int next_g) {
  const PmtModelDesc& D = P.d;
  const PmtBlockOffsets& BO = D.blocks[blk];
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  const int H = D.d_ffn / 2;
  const int g2 = P.blk_g0 + 2 * blk + 1;
  const float* W = C.W;
  const bool is_alt = row >= C.M->ref_pad;
  const int my_var = C.M->rowvar[row];
  const float alpha = __ldg(W + (is_alt ? BO.alpha_alt : BO.alpha_ref));
  const float beta = __ldg(W + (is_alt ? BO.beta_alt : BO.beta_ref));
  const float gamma = __ldg(W + BO.gamma);
  const int f_lo = part_lo(H, part), f_hi = part_lo(H, part + 1);
  for (int f = f_lo; f < f_hi; ++f)
    C.T1[f * LD + row] = C.T2[f * LD + row] * gate_value(P, C, BO, row, f, C.T2[(H + f) * LD + row], my_var, is_alt, alpha, beta, gamma);
  const float* img2 = stage.acquire(g2);
  stage.prefetch(next_g);
  gemm_tile(C.T1, P.gemm[g2], img2, W, C.M->ref_pad, C.X, EPI_RESIDUAL, 1.f, C.rows_used);
  __syncthreads();
}

// Picks the two buffers that are not `used`
__device__ __forceinline__ void other_two(const TileCtx& C, const float* used, float*& a, float*& b) {
  if (used == C.X) { a = C.T1; b = C.T2; }
  else if (used == C.T1) { a = C.T2; b = C.X; }
  else { a = C.T1; b = C.X; }
}

// euclidean_transformation.py:19-20: Fb = Q (y + t) per row
__device__ __forceinline__ void tile_rotate(const PmtModelDesc& D, const float* W, const float* y, float* Fb) {
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  const int E = D.d_feat;
  const int e_lo = part_lo(E, part), e_hi = part_lo(E, part + 1);
  for (int i = e_lo; i < e_hi; ++i) {
    float acc = 0.f;
    for (int j = 0; j < E; ++j) acc = fmaf(__ldg(W + D.rotation + i * E + j), y[j * LD + row] + __ldg(W + D.translation + j), acc);
    Fb[i * LD + row] = acc;
  }
}

// feature_clustering.py:82-119: per alt read K+2 log-likelihoods into Lb[0..K+2)
__device__ __forceinline__ void tile_head(const PmtModelDesc& D, const float* W, const HeadConst* HC, const TileMeta& M,
                                          const float* Fb, float* Lb) {
  const int row = threadIdx.x & (TILE - 1), part = threadIdx.x / TILE;
  const int E = D.d_feat, K = D.n_clusters;
  if (row >= M.ref_pad && M.rowvar[row] >= 0) {
    if (part == 0) {
      float q = 0.f, q2 = 0.f;
      for (int e = 0; e < E; ++e) {
        const float x = Fb[e * LD + row];
        const float a = x / HC->sigma[e], b = x / (2.f * HC->sigma[e]);
        q = fmaf(a, a, q); q2 = fmaf(b, b, q2);
      }
      Lb[0 * LD + row] = HC->c_non - q / 2.f;
      Lb[1 * LD + row] = HC->c_out - q2 / 2.f;
    }
    for (int k = part; k < K; k += NPART) {
      const float* u = W + D.unit_ke + k * E;
      float p = 0.f;
      for (int e = 0; e < E; ++e) p = fmaf(Fb[e * LD + row], __ldg(u + e), p);
      float o2 = 0.f;
      for (int e = 0; e < E; ++e) { const float d = Fb[e * LD + row] - p * __ldg(u + e); o2 = fmaf(d, d, o2); }
      const float dist = sqrtf(o2);
      const float ll_orth = HC->c_orth[k] - (dist * dist) / HC->two_tau2[k];
      const float ll_par = HC->log_half_lambda[k] + logerfc((HC->shift[k] - p) / HC->sqrt2_sigma[k]) +
                           HC->half_lambda[k] * (HC->two_mu_plus[k] - 2.f * p);
      Lb[(2 + k) * LD + row] = ll_orth + ll_par;
    }
  }
}

// Per-variant outputs from the accumulated sums (artifact_model.py:290-297, feature_clustering.py:121-135, :62-73)
__device__ __forceinline__ void tile_outputs(const Plan& P, const TileCtx& C, const PmtOutputs& out) {
  const PmtModelDesc& D = P.d;
  const int E = D.d_feat, K = D.n_clusters;
  const TileMeta& M = *C.M;
  for (int idx = threadIdx.x; idx < M.nv * E; idx += NTHREADS) {
    const int j = idx / E, e = idx % E;
    const long long v = M.v0 + j;
    if (out.alt_means_be) out.alt_means_be[v * E + e] = C.sums[(j * 2 + 1) * P.sum_w + e] / (M.alt_total[j] + 1e-4f);
    if (out.ref_means_be) out.ref_means_be[v * E + e] = C.sums[(j * 2 + 0) * P.sum_w + e] / (M.ref_total[j] + 1e-4f);
  }
  if (threadIdx.x < M.nv) {
    const int j = threadIdx.x;
    const long long v = M.v0 + j;
    const float* ll = C.llsum + (j * 2 + 1) * 16;
    const float non = ll[0], outl = ll[1];
    float art_max = -INFINITY;
    for (int k = 0; k < K; ++k) art_max = fmaxf(art_max, ll[2 + k] + C.HC->logw[k]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf(ll[2 + k] + C.HC->logw[k] - art_max);
    const float art = art_max + logf(s);
    if (out.logits_bk) {
      out.logits_bk[v * (K + 2) + 0] = non;
      out.logits_bk[v * (K + 2) + 1] = outl;
      for (int k = 0; k < K; ++k) out.logits_bk[v * (K + 2) + 2 + k] = ll[2 + k] + C.HC->logw[k];
    }
    if (out.logits_b) out.logits_b[v] = 20.f * tanhf((art - non) / 20.f);
    if (out.outlier_logits_b) out.outlier_logits_b[v] = outl - logsumexp2(non, art);
  }
}

// reducer -> rotation -> head -> per-variant sums (accumulating when the variant spans several chunks).
// Returns through Fb the buffer holding the final per-read features; `y_out` receives the reducer output buffer.
__device__ __forceinline__ void tile_tail(const Plan& P, TileCtx& C, Stage& stage, const PmtOutputs& out, bool accumulate,
                                          float* scr, float*& y_out, float*& Fb_out) {
  const PmtModelDesc& D = P.d;
  const int E = D.d_feat, K = D.n_clusters;
  float* red = run_mlp(P, D.red_ops, D.n_red_ops, P.red_g0, C.X, C.X, C.T1, C.T2, stage, C.W, C.rows_used, -1, scr, P.scr_red);
  __syncthreads();
  if (scr) save_rows(red, E, scr + P.scr_red[D.n_red_ops]);
  float *Fb, *Lb;
  other_two(C, red, Fb, Lb);
  tile_rotate(D, C.W, red, Fb);
  __syncthreads();
  tile_head(D, C.W, C.HC, *C.M, Fb, Lb);
  __syncthreads();
  segment_sums(*C.M, Fb, 0, E, C.sums, P.sum_w, false, accumulate);
  segment_sums(*C.M, Lb, 0, K + 2, C.llsum, 16, true, accumulate);
  if (out.final_re) {
    for (int idx = threadIdx.x; idx < TILE * E; idx += NTHREADS) {
      const int r = idx / E, e = idx % E;
      const long long n = C.M->rowidx[r];
      if (n >= 0) out.final_re[n * E + e] = Fb[e * LD + r];
    }
  }
  __syncthreads();
  y_out = red;
  Fb_out = Fb;
}

// Shared-memory carve-up common to the read-path kernels.  Returns the first free float after it.
__device__ __forceinline__ float* carve_tile_ctx(const Plan& P, float* smem, TileCtx& C, float*& st0, float*& st1, int stage_floats) {
  C.X = smem;
  C.T1 = C.X + PMT_MAX_DIM * LD;
  C.T2 = C.T1 + PMT_MAX_DIM * LD;
  st0 = C.T2 + PMT_MAX_DIM * LD;
  st1 = st0 + stage_floats;
  C.sums = st1 + stage_floats;
  C.llsum = C.sums + TILE * 2 * P.sum_w;
  C.HC = reinterpret_cast<HeadConst*>(C.llsum + TILE * 2 * 16);
  C.M = reinterpret_cast<TileMeta*>((reinterpret_cast<uintptr_t>(C.HC + 1) + 15) & ~uintptr_t(15));
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(C.M + 1) + 15) & ~uintptr_t(15));
}
__host__ __device__ inline size_t tile_ctx_bytes(const Plan& P, int stage_floats) {
  return (size_t)(3 * PMT_MAX_DIM * LD + 2 * stage_floats + TILE * 2 * P.sum_w + TILE * 2 * 16) * sizeof(float) +
         sizeof(HeadConst) + sizeof(TileMeta) + 64;
}

}  // namespace pmt
