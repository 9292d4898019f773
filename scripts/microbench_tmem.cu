// TMEM as a per-thread row store: can 8 warps keep a 50257-word row in tensor memory (tcgen05.st / tcgen05.ld,
// 32x32b.x4 = one float4 per thread per instruction) and sweep it as fast as 8 other warps sweep a row in shared
// memory?  Sweeps: an fp64-exp pass in place (the coder's P1) and a light pass (LDS/LDTM + 4 adds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_bin/mb_tmem scripts/microbench_tmem.cu
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

__device__ __forceinline__ double exp_like(double a, const double* tab) {
  const double magic = 6755399441055744.0;
  double t = __fma_rn(a, 738.6598609246875, magic);
  int n = (int)(uint32_t)__double_as_longlong(t);
  double nd = t - magic;
  double r = __fma_rn(nd, -0.0013537890625, a);
  r = __fma_rn(nd, -1.1e-13, r);
  double T = tab[n & 511];
  double q = __fma_rn(r, 1.0 / 24.0, 1.0 / 6.0);
  q = __fma_rn(q, r, 0.5);
  double r2 = r * r;
  double p = __fma_rn(q, r2, r);
  double e = __fma_rn(T, p, T);
  const int hi = __double2hiint(e) + ((n & ~511) << 11);
  return __hiloint2double(hi, __double2loint(e));
}
__device__ __forceinline__ float pack_e(double e) {
  return __uint_as_float(__funnelshift_l((uint32_t)__double2loint(e), (uint32_t)__double2hiint(e), 4));
}

// mode bit 0: group A (warps 0-7) sweeps a shared-memory row; bit 1: group B (warps 8-15) sweeps a TMEM row
// heavy: fp64 exp pass in place; else light pass
__global__ void __launch_bounds__(512, 1) k_sweep(int mode, int heavy, int reps, int nchunk, u64* out, float* sink, int* bad) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t tbase;
  __shared__ double tab[512];
  float4* row = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tab[tid] = 1.0 + tid * (1.0 / 1024.0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase;
  const int g = warp >> 3;                 // group
  const int gt = tid & 255;                // thread within group
  // TMEM address of this thread's chunk j: lanes of the warp's quadrant, columns (half of the quadrant's 512) + 4j
  const uint32_t my_t = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * 256);
  const int per = (nchunk + 255) / 256;    // chunks per thread
  // fill
  if (g == 0) {
    for (int c = gt; c < nchunk; c += 256) row[c] = make_float4(-0.001f * c, -0.002f * c, -0.0005f * c, -0.0001f * (c & 1023));
  } else {
    for (int j = 0; j < per; ++j) {
      const int c = gt + j * 256;
      tm_st4(my_t + 4 * j, make_float4(-0.001f * c, -0.002f * c, -0.0005f * c, -0.0001f * (c & 1023)));
    }
    tm_wait_st();
    // verify
    int nb = 0;
    for (int j = 0; j < per; ++j) {
      const int c = gt + j * 256;
      float4 v = tm_ld4(my_t + 4 * j);
      tm_wait_ld();
      if (v.x != -0.001f * c || v.y != -0.002f * c || v.z != -0.0005f * c || v.w != -0.0001f * (c & 1023)) ++nb;
    }
    if (nb) atomicAdd(bad, nb);
  }
  __syncthreads();
  double acc = 0.0;
  float facc = 0.f;
  const long long t0 = clock64();
  if (g == 0 && (mode & 1)) {
    for (int r = 0; r < reps; ++r) {
      if (heavy) {
        for (int c = gt; c < nchunk; c += 256) {
          const float4 v = row[c];
          const double e0 = exp_like((double)v.x, tab), e1 = exp_like((double)v.y, tab), e2 = exp_like((double)v.z, tab), e3 = exp_like((double)v.w, tab);
          acc += (e0 + e1) + (e2 + e3);
          row[c] = make_float4(pack_e(e0), pack_e(e1), pack_e(e2), pack_e(e3));
        }
      } else {
        for (int c = gt; c < nchunk; c += 512) {
          const float4 v = row[c];
          const float4 w = (c + 256 < nchunk) ? row[c + 256] : make_float4(0.f, 0.f, 0.f, 0.f);
          facc += (v.x + v.y) + (v.z + v.w) + (w.x + w.y) + (w.z + w.w);
        }
      }
      asm volatile("bar.sync 1, 256;");
    }
  }
  if (g == 1 && (mode & 2)) {
    for (int r = 0; r < reps; ++r) {
      if (heavy) {
        for (int j = 0; j < per; ++j) {
          const float4 v = tm_ld4(my_t + 4 * j);
          tm_wait_ld();
          const double e0 = exp_like((double)v.x, tab), e1 = exp_like((double)v.y, tab), e2 = exp_like((double)v.z, tab), e3 = exp_like((double)v.w, tab);
          acc += (e0 + e1) + (e2 + e3);
          tm_st4(my_t + 4 * j, make_float4(pack_e(e0), pack_e(e1), pack_e(e2), pack_e(e3)));
        }
        tm_wait_st();
      } else {
        for (int j = 0; j < per; j += 2) {
          const float4 v = tm_ld4(my_t + 4 * j);
          const float4 w = tm_ld4(my_t + 4 * (j + 1 < per ? j + 1 : j));
          tm_wait_ld();
          facc += (v.x + v.y) + (v.z + v.w) + (w.x + w.y) + (w.z + w.w);
        }
      }
      asm volatile("bar.sync 2, 256;");
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = (u64)(t1 - t0);
  if (acc == 123.456 || facc == 123.456f) sink[0] = (float)acc + facc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tb));
}

int main() {
  const int nchunk = 12565, reps = 20;
  u64* out; float* sink; int* bad;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4); cudaMalloc(&bad, 4);
  cudaMemset(bad, 0, 4);
  cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  for (int heavy = 1; heavy >= 0; --heavy)
    for (int mode = 1; mode <= 3; ++mode) {
      k_sweep<<<148, 512, 210 * 1024>>>(mode, heavy, reps, nchunk, out, sink, bad);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      u64 h[148 * 16];
      int hb;
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
      double a = 0, b = 0;
      for (int i = 0; i < 148; ++i) { a += (double)h[i * 16 + 0]; b += (double)h[i * 16 + 8]; }
      printf("%s pass, mode %d (1=smem group, 2=tmem group, 3=both): cycles per row-sweep  smem %.0f  tmem %.0f   (tmem readback mismatches %d)\n",
             heavy ? "fp64-exp" : "light", mode, a / 148 / reps, b / 148 / reps, hb);
    }
  return 0;
}
