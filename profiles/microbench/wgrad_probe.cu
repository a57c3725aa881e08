// Hardware probe for the tensor-core backward (sm_100a).  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o wgrad_probe wgrad_probe.cu && ./wgrad_probe
// One shared-memory copy of a per-row tile X[128 rows][W features], stored as 16-byte chunks
//   addr(r, c) = (r / 8) * GROUP + c * 128 + (r % 8) * 16          (8 rows x 16 B = one 128-byte core matrix)
// is read by tcgen05.mma (kind::tf32, no swizzle) in two ways:
//   (1) wgrad:  D[m][n] = sum_r ACT[r][m] * DY[r][n]   both operands MN-major (the reduction runs over the rows),
//               16 MMAs of K = 8 rows, M = 128 (rows m >= W of D are garbage and ignored)
//   (2) dgrad:  D[r][k] = sum_n DY[r][n] * WT[k][n]     A = DY K-major from the same buffer, B = 128B-swizzled image
// and the probe checks both against the CPU and times them.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 128B-swizzled K-major image (weights)
__device__ __forceinline__ uint64_t desc_sw128(unsigned addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// no swizzle: leading / stride byte offsets explicit
__device__ __forceinline__ uint64_t desc_plain(unsigned addr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// SWIZZLE_128B_BASE32B (layout type 1): the only layout of MN-major tf32 operands
__device__ __forceinline__ uint64_t desc_b32(unsigned addr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ unsigned make_idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(unsigned d, uint64_t a, uint64_t b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
#define LD16(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr))
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

struct Sh {
  unsigned long long bar_a, bar_d;
  unsigned tmem_base;
};

constexpr int DY_GROUP = 2048;
constexpr int PANEL = 16384;   // one panel: 128 rows x 32 features (128 B per row), 32-byte units XOR-swizzled by row % 4   // bytes between 8-row groups of the dY buffer (16 chunks)

// mode 0: wgrad, mode 1: dgrad.  act: [128][WA], dy: [128][WN], img: swizzled [WA x WN] transposed weights (dgrad).
__global__ void __launch_bounds__(160, 1) probe(const float* __restrict__ act, const float* __restrict__ dy, const float* __restrict__ img,
                                                float* __restrict__ D, int WA, int WN, int mode, int reps, long long* cycles, int variant) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* act_s = p; p += 32 * 1024;
  unsigned char* dy_s = p; p += 32 * 1024;
  unsigned char* img_s = p; p += 32 * 1024;
  Sh* S = reinterpret_cast<Sh*>(p);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_addr(&S->bar_a), 128); mbar_init(smem_addr(&S->bar_d), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const int n_kb = (WN + 31) / 32;
  for (int i = tid; i < n_kb * WA * 32; i += blockDim.x) reinterpret_cast<float*>(img_s)[i] = img[i];
  for (int i = tid; i < 32 * 1024 / 4; i += blockDim.x) { reinterpret_cast<float*>(act_s)[i] = 0.f; reinterpret_cast<float*>(dy_s)[i] = 0.f; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = S->tmem_base;
  const int act_group = (WA / 4) * 128;
  if (warp < 4) {
    const int r = tid;
    for (int c = 0; c < WA / 4; ++c)
      *reinterpret_cast<float4*>(act_s + (c >> 3) * PANEL + r * 128 + ((((c & 7) >> 1) ^ (r & 3)) << 5) + (c & 1) * 16) = *reinterpret_cast<const float4*>(act + r * WA + c * 4);
    for (int c = 0; c < WN / 4; ++c)
      *reinterpret_cast<float4*>(dy_s + (c >> 3) * PANEL + r * 128 + ((((c & 7) >> 1) ^ (r & 3)) << 5) + (c & 1) * 16) = *reinterpret_cast<const float4*>(dy + r * WN + c * 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    mbar_arrive(smem_addr(&S->bar_a));
    mbar_wait(smem_addr(&S->bar_d), 0);
    tc_fence_after();
    const unsigned trow = tm + ((unsigned)(warp * 32) << 16);
    const int ncol = mode == 0 ? WN : WA;
    unsigned v[16];
    for (int c0 = 0; c0 < ncol; c0 += 16) {
      LD16(trow + c0, v);
      WAIT_LD();
      for (int j = 0; j < 16; ++j) D[r * ncol + c0 + j] = __uint_as_float(v[j]);
    }
  } else if (tid == 128) {
    mbar_wait(smem_addr(&S->bar_a), 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      if (mode == 0) {
        const unsigned idesc = make_idesc(WN, 1, 1);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          if (variant == 0)
            mma_ss(tm, desc_b32(smem_addr(act_s) + ks * 1024, PANEL, 512), desc_b32(smem_addr(dy_s) + ks * 1024, PANEL, 512), idesc, ks > 0);
          else
            mma_ss(tm, desc_b32(smem_addr(act_s) + ks * 1024, 512, PANEL), desc_b32(smem_addr(dy_s) + ks * 1024, 512, PANEL), idesc, ks > 0);
      } else if (mode == 2) {
        const unsigned idesc = make_idesc(WA, 0, 1);
        for (int ks = 0; ks < WN / 8; ++ks) {
          const uint64_t ad = desc_b32(smem_addr(dy_s) + (ks >> 2) * PANEL + (ks & 3) * 32, 16, 1024);
          if (variant == 0) mma_ss(tm, ad, desc_b32(smem_addr(act_s) + ks * 1024, PANEL, 512), idesc, ks > 0);
          else mma_ss(tm, ad, desc_b32(smem_addr(act_s) + ks * 1024, 512, PANEL), idesc, ks > 0);
        }
      } else {
        const unsigned idesc = make_idesc(WA, 0, 0);
        for (int ks = 0; ks < WN / 8; ++ks) {
          const unsigned boff = (ks >> 2) * (WA * 128) + (ks & 3) * 32;
          mma_ss(tm, desc_b32(smem_addr(dy_s) + (ks >> 2) * PANEL + (ks & 3) * 32, 16, 1024), desc_sw128(smem_addr(img_s) + boff), idesc, ks > 0);
        }
      }
    }
    mma_commit(smem_addr(&S->bar_d));
    mbar_wait(smem_addr(&S->bar_d), 0);
    if (cycles) *cycles = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }


int main() {
  const int shapes[][2] = {{64, 64}, {64, 48}, {32, 32}, {24, 64}, {64, 16}};   // (WA = operand width, WN = dY width)
  float *d_act, *d_dy, *d_img, *d_D;
  long long* d_cyc;
  CK(cudaMalloc(&d_act, 128 * 64 * 4)); CK(cudaMalloc(&d_dy, 128 * 64 * 4)); CK(cudaMalloc(&d_img, 64 * 64 * 4 * 2)); CK(cudaMalloc(&d_D, 128 * 64 * 4));
  CK(cudaMalloc(&d_cyc, 8));
  const int smem = 3 * 32 * 1024 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  {   // layout diagnostic: ACT one-hot at (r0, m0), DY[r][n] = 100 r + n  ->  D[m][n] = [m == m0] * (100 r0 + n) if the layout is as assumed
    const int WA = 64, WN = 64;
    std::vector<float> act(128 * WA), dy(128 * WN), img(2 * WA * 32, 0.f), D(128 * 64);
    for (int r = 0; r < 128; ++r) for (int n = 0; n < WN; ++n) dy[r * WN + n] = 100.f * r + n;
    CK(cudaMemcpy(d_dy, dy.data(), dy.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    const int pts[][2] = {{0, 0}, {1, 0}, {0, 1}, {0, 4}, {3, 5}, {8, 0}, {9, 17}, {2, 33}, {127, 63}};
    for (int variant = 0; variant < 2; ++variant)
      for (auto& pt : pts) {
        std::fill(act.begin(), act.end(), 0.f);
        act[pt[0] * WA + pt[1]] = 1.f;
        CK(cudaMemcpy(d_act, act.data(), act.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_D, 0, 128 * 64 * 4));
        probe<<<1, 160, smem>>>(d_act, d_dy, d_img, d_D, WA, WN, 0, 1, nullptr, variant);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), d_D, 128 * WN * 4, cudaMemcpyDeviceToHost));
        printf("v%d act(r=%d,m=%d):", variant, pt[0], pt[1]);
        int shown = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < WN; ++n)
            if (D[m * WN + n] != 0.f && shown < 6) { printf("  D[%d][%d]=%.0f", m, n, D[m * WN + n]); ++shown; }
        printf("\n");
      }
  }
  for (auto& sh : shapes) {
    const int WA = sh[0], WN = sh[1];
    std::vector<float> act(128 * WA), dy(128 * WN), wt(WA * WN), img(((WN + 31) / 32) * WA * 32, 0.f), D(128 * 64);
    srand(WA * 131 + WN);
    for (auto& v : act) v = tf32_trunc((rand() % 2001 - 1000) / 500.f);
    for (auto& v : dy) v = tf32_trunc((rand() % 2001 - 1000) / 500.f);
    for (auto& v : wt) v = tf32_trunc((rand() % 2001 - 1000) / 500.f);   // WT[k][n]: k < WA (output), n < WN (reduction)
    for (int k = 0; k < WA; ++k)
      for (int n = 0; n < WN; ++n) {
        const int kb = n / 32, kk = n % 32;
        const unsigned L = (unsigned)k * 128u + (unsigned)kk * 4u;
        const unsigned phys = L ^ (((L >> 7) & 7u) << 4);
        img[(size_t)kb * WA * 32 + phys / 4] = wt[k * WN + n];
      }
    CK(cudaMemcpy(d_act, act.data(), act.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_dy, dy.data(), dy.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    for (int mv = 0; mv < 5; ++mv) {
      const int mode = mv == 2 ? 1 : (mv >= 3 ? 2 : 0), variant = mv >= 3 ? mv - 3 : mv;
      if (mode == 2) {   // the act buffer holds X[n][k] = WT[k][n]
        std::vector<float> xt(128 * WA, 0.f);
        for (int n = 0; n < WN; ++n) for (int k = 0; k < WA; ++k) xt[n * WA + k] = wt[k * WN + n];
        CK(cudaMemcpy(d_act, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice));
      }
      if (mode == 0 && WN % 16) continue;
      if (mode >= 1 && WA % 16) continue;
      CK(cudaMemset(d_D, 0, 128 * 64 * 4));
      probe<<<1, 160, smem>>>(d_act, d_dy, d_img, d_D, WA, WN, mode, 1, nullptr, variant);
      CK(cudaDeviceSynchronize());
      const int ncol = mode == 0 ? WN : WA;
      CK(cudaMemcpy(D.data(), d_D, 128 * ncol * 4, cudaMemcpyDeviceToHost));
      double worst = 0;
      if (mode == 0) {
        for (int m = 0; m < WA; ++m)
          for (int n = 0; n < WN; ++n) {
            double s = 0;
            for (int r = 0; r < 128; ++r) s += (double)act[r * WA + m] * dy[r * WN + n];
            worst = fmax(worst, fabs(s - D[m * WN + n]));
          }
      } else {
        for (int r = 0; r < 128; ++r)
          for (int k = 0; k < WA; ++k) {
            double s = 0;
            for (int n = 0; n < WN; ++n) s += (double)dy[r * WN + n] * wt[k * WN + n];
            worst = fmax(worst, fabs(s - D[r * WA + k]));
          }
      }
      probe<<<1, 160, smem>>>(d_act, d_dy, d_img, d_D, WA, WN, mode, 64, d_cyc, variant);
      CK(cudaDeviceSynchronize());
      long long cyc = 0;
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      const int n_mma = mode == 0 ? 16 : WN / 8;
      printf("%s v%d WA=%d WN=%d  max|err|=%.3e  %s   %.1f cycles per MMA (%d MMAs per layer, %.0f cycles per layer)\n", mode == 0 ? "wgrad" : (mode == 1 ? "dgrad" : "dgrad-Bmn"), variant, WA, WN,
             worst, worst < 1e-3 ? "OK" : "WRONG", (double)cyc / (64.0 * n_mma), n_mma, (double)cyc / 64.0);
    }
  }
  return 0;
}
