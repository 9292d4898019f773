// Micro-benchmarks that calibrate the coder kernel's cost model on B200 (cycles per warp instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench scripts/microbench.cu && ./microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ACTIVE>   // lanes per warp that really add (others predicated off)
__global__ void k_red(unsigned long long* out, int iters, uint32_t seed) {
  __shared__ uint32_t hist[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  uint32_t x = seed ^ (threadIdx.x * 2654435761u);
  const bool on = (threadIdx.x & 31) < ACTIVE;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    const uint32_t bin = (x >> 11) & 2047u;
    const uint32_t q = on ? 1u : 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" :: "r"(smem_addr(hist + bin)), "r"(q) : "memory");
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (hist[5] & 1);
}

__global__ void k_sts(unsigned long long* out, int iters, uint32_t seed) {
  __shared__ uint32_t hist[2048];
  uint32_t x = seed ^ (threadIdx.x * 2654435761u);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    hist[(x >> 11) & 2047u] = x;
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (hist[5] & 1);
}

template <int CHAINS>
__global__ void k_dfma(unsigned long long* out, int iters, double a, double b) {
  double v[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) v[j] = a + j + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) v[j] = __fma_rn(v[j], b, a);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += v[j];
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (s == 12345.0);
}

template <int CHAINS>
__global__ void k_mufu(unsigned long long* out, int iters, float a) {
  float v[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) v[j] = a + j * 0.001f + threadIdx.x * 1e-6f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[j])); v[j] = y - 1.0f; }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += v[j];
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (s == 12345.0f);
}

template <int CHAINS>
__global__ void k_f2f(unsigned long long* out, int iters, float a) {
  float v[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) v[j] = a + j + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) { double d = (double)v[j]; v[j] = __double2float_rz(d * 1.0000001); }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += v[j];
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (s == 12345.0f);
}

// 8 DFMA chains + NI independent integer multiply-adds per iteration: does the fp64 issue cost add to the
// other instructions' (2 issue cycles per DFMA) or overlap with them?
template <int NI>
__global__ void k_mix(unsigned long long* out, int iters, double a, double b, uint32_t m) {
  double v[8];
  uint32_t u[NI > 0 ? NI : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = a + j + threadIdx.x;
#pragma unroll
  for (int j = 0; j < (NI > 0 ? NI : 1); ++j) u[j] = threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __fma_rn(v[j], b, a);
#pragma unroll
    for (int j = 0; j < NI; ++j) u[j] = (u[j] ^ m) + (u[j] >> 3);
  }
  long long t1 = clock64();
  double s = 0;
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
  for (int j = 0; j < (NI > 0 ? NI : 1); ++j) x += u[j];
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (s == 12345.0) + (x == 77u);
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 148 * 8);
  unsigned long long h[148];
  const int it = 4096, T = 512, W = T / 32;
  auto report = [&](const char* name, double instr_per_thread_iter) {
    cudaDeviceSynchronize(); cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h[0];
    printf("%-34s %9.0f cycles  -> %.2f cycles per warp-instruction per SM (%.1f lanes/clk/SM)\n", name, cyc,
           cyc / (it * instr_per_thread_iter * W), it * instr_per_thread_iter * T / cyc);
  };
  k_red<32><<<148, T>>>(d, it, 1); report("red.shared 32 lanes random", 1);
  k_red<4><<<148, T>>>(d, it, 1); report("red.shared 4 lanes random", 1);
  k_red<1><<<148, T>>>(d, it, 1); report("red.shared 1 lane", 1);
  k_red<0><<<148, T>>>(d, it, 1); report("red.shared all predicated off", 1);
  k_sts<<<148, T>>>(d, it, 1); report("st.shared scatter random", 1);
  k_dfma<1><<<148, T>>>(d, it, 1.0, 0.999); report("DFMA 1 chain/thread", 1);
  k_dfma<4><<<148, T>>>(d, it, 1.0, 0.999); report("DFMA 4 chains/thread", 4);
  k_dfma<8><<<148, T>>>(d, it, 1.0, 0.999); report("DFMA 8 chains/thread", 8);
  k_mufu<4><<<148, T>>>(d, it, 0.5f); report("MUFU.EX2 (+FADD) 4 chains", 4);
  k_f2f<4><<<148, T>>>(d, it, 0.5f); report("F2F.F64.F32 + DMUL + F2F.F32.F64", 4);
  k_mix<0><<<148, T>>>(d, it, 1.0, 0.999, 5u); report("8 DFMA + 0 int pairs", 8);
  k_mix<4><<<148, T>>>(d, it, 1.0, 0.999, 5u); report("8 DFMA + 4 int pairs (8 instr)", 8);
  k_mix<8><<<148, T>>>(d, it, 1.0, 0.999, 5u); report("8 DFMA + 8 int pairs (16 instr)", 8);
  k_mix<16><<<148, T>>>(d, it, 1.0, 0.999, 5u); report("8 DFMA + 16 int pairs (32 instr)", 8);
  return 0;
}
