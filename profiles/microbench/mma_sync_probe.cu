// Throughput probe for the legacy warp-level tensor path on sm_100a: mma.sync.m16n8k8 (tf32) and m16n8k16 (bf16),
// 8 / 16 warps per SM, four independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_probe mma_sync_probe.cu && ./mma_sync_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void probe(long long* out, float* sink, int iters) {
  float c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3e800000u};
  unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float s = 0.f;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  for (int kind = 0; kind < 2; ++kind)
    for (int warps : {4, 8, 16, 32}) {
      long long h = 0;
      if (kind == 0) probe<0><<<148, warps * 32>>>(d_out, sink, iters); else probe<1><<<148, warps * 32>>>(d_out, sink, iters);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      const double mmas = (double)iters * 4 * warps;             // per SM
      const double flop = mmas * 2.0 * 16 * 8 * (kind == 0 ? 8 : 16);
      printf("%s warps/SM=%2d: %.2f cycles per MMA per SM, %.0f FLOP/cycle/SM -> %.0f TFLOP/s at 1.965 GHz x 148\n",
             kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", warps, h / mmas, flop / h, flop / h * 148 * 1.965e9 / 1e12);
    }
  return 0;
}
