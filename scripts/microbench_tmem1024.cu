// Does a row kept in tensor memory (thread-private, tcgen05.ld / tcgen05.st 32x32b.x4) relieve the shared-memory pipe of the
// exp pass?  One 1024-thread CTA per SM as in ns_lean.cuh; sweeps of a 50257-word row: the fp64-exp pass in place (P1, with
// its random 8-byte table lookups in shared memory) and a light read pass, with the row (a) in shared memory, (b) in TMEM
// (32 warps: 8 per lane quadrant, 64 columns each = 13 float4 chunks per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_bin/mb_tmem1024 scripts/microbench_tmem1024.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef unsigned long long u64;

__device__ __forceinline__ void tm_st4(uint32_t taddr, float4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ float4 tm_ld4(uint32_t taddr) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
  return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool FREE>
__device__ __forceinline__ double exp_like_t(double a, const double* tab) {
  const double magic = 6755399441055744.0;
  double t = __fma_rn(a, 738.6598609246875, magic);
  int n = (int)(uint32_t)__double_as_longlong(t);
  double nd = (double)n;
  double r = __fma_rn(nd, -0.0013537890625, a);
  r = __fma_rn(nd, -1.1e-13, r);
  double T = FREE ? tab[(threadIdx.x & 15) + ((n >> 30) & 1)] : tab[n & 511];   // FREE: one bank pair per lane of a half-warp (timing only)
  double q = __fma_rn(r, 1.0 / 24.0, 1.0 / 6.0);
  q = __fma_rn(q, r, 0.5);
  double r2 = r * r;
  double p = __fma_rn(q, r2, r);
  double e = __fma_rn(T, p, T);
  const int hi = __double2hiint(e) + ((n & ~511) << 11);
  return __hiloint2double(hi, __double2loint(e));
}
__device__ __forceinline__ double exp_like(double a, const double* tab) { return exp_like_t<false>(a, tab); }
__device__ __forceinline__ float pack_e(double e) {
  return __uint_as_float(__funnelshift_l((uint32_t)__double2loint(e), (uint32_t)__double2hiint(e), 4));
}
__device__ __forceinline__ float4 fill_of(int c) {   // pseudo-random logits in [-12, 0]
  uint32_t h = (uint32_t)c * 2654435761u;
  auto f = [&](uint32_t k) { h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; return -12.0f * (float)((h + k) & 0xffffff) / 16777216.0f; };
  return make_float4(f(1), f(2), f(3), f(4));
}

__global__ void __launch_bounds__(1024, 1) k_sweep(int tmem, int heavy, int reps, int nchunk, u64* out, float* sink, int* bad) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t tbase;
  __shared__ double tab[512];
  float4* row = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 512) tab[tid] = 1.0 + tid * (1.0 / 1024.0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase;
  const uint32_t my_t = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  const int per = (nchunk + 1023) / 1024;   // 13
  for (int c = tid; c < nchunk; c += 1024) row[c] = fill_of(c);
  for (int j = 0; j < per; ++j) tm_st4(my_t + 4 * j, fill_of(tid + j * 1024));
  tm_wait_st();
  int nb = 0;
  for (int j = 0; j < per; ++j) {
    float4 v = tm_ld4(my_t + 4 * j);
    tm_wait_ld();
    const float4 w = fill_of(tid + j * 1024);
    if (v.x != w.x || v.y != w.y || v.z != w.z || v.w != w.w) ++nb;
  }
  if (nb) atomicAdd(bad, nb);
  __syncthreads();
  double acc = 0.0;
  float facc = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (tmem == 2) {
#pragma unroll 1
      for (int c = tid; c < nchunk; c += 1024) {
        const float4 v = row[c];
        const double e0 = exp_like_t<true>((double)v.x, tab), e1 = exp_like_t<true>((double)v.y, tab), e2 = exp_like_t<true>((double)v.z, tab), e3 = exp_like_t<true>((double)v.w, tab);
        acc += (e0 + e1) + (e2 + e3);
        row[c] = make_float4(-pack_e(e0) * 1e-30f, v.y, v.z, v.w);
      }
    } else if (!tmem) {
      if (heavy) {
#pragma unroll 1
        for (int c = tid; c < nchunk; c += 1024) {
          const float4 v = row[c];
          const double e0 = exp_like((double)v.x, tab), e1 = exp_like((double)v.y, tab), e2 = exp_like((double)v.z, tab), e3 = exp_like((double)v.w, tab);
          acc += (e0 + e1) + (e2 + e3);
          row[c] = make_float4(-pack_e(e0) * 1e-30f, v.y, v.z, v.w);   // keep the inputs in range for the next repetition
        }
      } else {
#pragma unroll 1
        for (int c = tid; c < nchunk; c += 1024) { const float4 v = row[c]; facc += (v.x + v.y) + (v.z + v.w); }
      }
    } else {
      if (heavy) {
        float4 v = tm_ld4(my_t);
#pragma unroll 1
        for (int j = 0; j < per; ++j) {
          tm_wait_ld();
          const float4 cur = v;
          if (j + 1 < per) v = tm_ld4(my_t + 4 * (j + 1));      // next chunk in flight
          const double e0 = exp_like((double)cur.x, tab), e1 = exp_like((double)cur.y, tab), e2 = exp_like((double)cur.z, tab), e3 = exp_like((double)cur.w, tab);
          acc += (e0 + e1) + (e2 + e3);
          tm_st4(my_t + 4 * j, make_float4(-pack_e(e0) * 1e-30f, cur.y, cur.z, cur.w));
        }
        tm_wait_st();
      } else {
#pragma unroll 1
        for (int j = 0; j < per; ++j) { const float4 v = tm_ld4(my_t + 4 * j); tm_wait_ld(); facc += (v.x + v.y) + (v.z + v.w); }
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 32 + warp] = (u64)(t1 - t0);
  if (acc == 123.456 || facc == 123.456f) sink[0] = (float)acc + facc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tb));
}

int main() {
  const int nchunk = 12565, reps = 20;
  u64* out; float* sink; int* bad;
  cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&sink, 4); cudaMalloc(&bad, 4);
  cudaMemset(bad, 0, 4);
  cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  for (int heavy = 1; heavy >= 0; --heavy)
    for (int tmem = 0; tmem <= (heavy ? 2 : 1); ++tmem) {
      k_sweep<<<148, 1024, 210 * 1024>>>(tmem, heavy, reps, nchunk, out, sink, bad);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      u64 h[148 * 32];
      int hb;
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
      double a = 0;
      for (int i = 0; i < 148; ++i) a += (double)h[i * 32];
      printf("%s pass, row in %s: %.0f cycles per row-sweep (tmem readback mismatches %d)\n", heavy ? "fp64-exp" : "light", tmem == 2 ? "shared memory, conflict-free table lookups (timing only)" : tmem ? "TMEM" : "shared memory", a / 148 / reps, hb);
    }
  return 0;
}
