// membench.cu — B200 micro-measurements that decide the hash-table layout.
// Not part of the product path. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
// Measures, with CUDA events on one stream:
//   stream   : int4 streaming read (sanity vs MEASURED_PEAKS.json)
//   gather4  : random 4-B loads  (direct-address table)       per table size
//   gather8  : random 8-B loads  (packed key|row slot)         per table size
//   gather32 : random 32-B bucket loads (LDG.256)              per table size
//   cas8     : random 8-B atomicCAS inserts                    per table size
//   smem     : random 8-B loads from a 128 KB shared-memory table
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol_first())); return r;
}
__device__ __forceinline__ void st_stream(int4* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol_first()));
}
__device__ __forceinline__ uint32_t ld4_keep(const uint32_t* p) {
  uint32_t r; asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol_last())); return r;
}
__device__ __forceinline__ uint64_t ld8_keep(const uint64_t* p) {
  uint64_t r; asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol_last())); return r;
}
__device__ __forceinline__ void ld32_keep(const uint64_t* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

__global__ void k_fill_keys(uint32_t* keys, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) keys[i] = mix32((uint32_t)i * 2654435761U + seed);
}

__global__ void k_stream(const int4* __restrict__ in, size_t n4, int* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  int acc = 0;
  for (; i < n4; i += stride) { int4 v = ld_stream(in + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678) *out = acc;
}

// mode 0: gather4, 1: gather8, 2: gather32. writeback: store 4 B/tuple result stream (like the match cache)
template <int MODE, bool WB>
__global__ void __launch_bounds__(256) k_gather(const int4* __restrict__ keys, size_t n4, const void* __restrict__ table, uint32_t mask, int4* __restrict__ res, int* sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n4; i += stride) {
    int4 k = ld_stream(keys + i);
    uint32_t h[4] = {(uint32_t)k.x & mask, (uint32_t)k.y & mask, (uint32_t)k.z & mask, (uint32_t)k.w & mask};
    uint32_t r[4];
    if (MODE == 0) {
      #pragma unroll
      for (int j = 0; j < 4; j++) r[j] = ld4_keep((const uint32_t*)table + h[j]);
    } else if (MODE == 1) {
      uint64_t v[4];
      #pragma unroll
      for (int j = 0; j < 4; j++) v[j] = ld8_keep((const uint64_t*)table + h[j]);
      #pragma unroll
      for (int j = 0; j < 4; j++) r[j] = (uint32_t)v[j] ^ (uint32_t)(v[j] >> 32);
    } else {
      uint64_t a[4], b[4], c[4], d[4];
      #pragma unroll
      for (int j = 0; j < 4; j++) ld32_keep((const uint64_t*)table + 4 * (size_t)h[j], a[j], b[j], c[j], d[j]);
      #pragma unroll
      for (int j = 0; j < 4; j++) { uint64_t x = a[j] ^ b[j] ^ c[j] ^ d[j]; r[j] = (uint32_t)x ^ (uint32_t)(x >> 32); }
    }
    if (WB) st_stream(res + i, make_int4(r[0], r[1], r[2], r[3]));
    else acc ^= r[0] ^ r[1] ^ r[2] ^ r[3];
  }
  if (!WB && acc == 0x12345678) *sink = acc;
}

__global__ void __launch_bounds__(256) k_cas8(const int4* __restrict__ keys, size_t n4, unsigned long long* table, uint32_t mask, int* sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n4; i += stride) {
    int4 k = ld_stream(keys + i);
    uint32_t kk[4] = {(uint32_t)k.x, (uint32_t)k.y, (uint32_t)k.z, (uint32_t)k.w};
    #pragma unroll
    for (int j = 0; j < 4; j++) {
      unsigned long long old = atomicCAS(table + (kk[j] & mask), ~0ULL, ((unsigned long long)kk[j] << 32) | (uint32_t)(4 * i + j));
      acc ^= (uint32_t)old;
    }
  }
  if (acc == 0x12345678) *sink = acc;
}

// shared-memory table gather: each CTA owns a 128 KB table (16384 x 8 B), streams keys, looks up.
__global__ void __launch_bounds__(512) k_smem(const int4* __restrict__ keys, size_t n4, int4* __restrict__ res, int* sink) {
  extern __shared__ uint64_t tab[];
  for (int t = threadIdx.x; t < 16384; t += blockDim.x) tab[t] = mix32(t) | ((uint64_t)t << 32);
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    int4 k = ld_stream(keys + i);
    uint64_t a = tab[(uint32_t)k.x & 16383], b = tab[(uint32_t)k.y & 16383], c = tab[(uint32_t)k.z & 16383], d = tab[(uint32_t)k.w & 16383];
    st_stream(res + i, make_int4((int)a, (int)b, (int)c, (int)d));
  }
}

template <class F>
static float best_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0) best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  size_t n = (size_t)1 << 28;                 // probe tuples
  if (argc > 1) n = (size_t)1 << atoi(argv[1]);
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"persist_l2_max\": %d, \"n\": %zu}\n", p.name, p.multiProcessorCount, p.l2CacheSize, p.persistingL2CacheMaxSize, n);
  uint32_t* keys; int4* res; void* table; int* sink;
  CK(cudaMalloc(&keys, n * 4)); CK(cudaMalloc(&res, n * 4)); CK(cudaMalloc(&table, (size_t)1 << 31)); CK(cudaMalloc(&sink, 4));
  k_fill_keys<<<148 * 8, 256>>>(keys, n, 12345u); CK(cudaDeviceSynchronize());
  CK(cudaMemset(table, 0xFF, (size_t)1 << 31));
  size_t n4 = n / 4;
  int grid = 148 * 8;
  {
    float ms = best_ms([&] { k_stream<<<grid, 256>>>((const int4*)keys, n4, sink); });
    printf("{\"bench\": \"stream_read\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, n * 4 / ms / 1e6);
  }
  for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    float ms = best_ms([&] { k_gather<1, true><<<g, 256>>>((const int4*)keys, n4, table, (1u << 22) - 1, res, sink); });
    printf("{\"bench\": \"gather8_wb_32MB_grid\", \"grid\": %d, \"ms\": %.4f, \"Glookups\": %.1f}\n", g, ms, n / ms / 1e6);
  }
  for (int lg = 21; lg <= 30; lg++) {          // table bytes = 2^lg ... 2 MB .. 1 GB
    size_t bytes = (size_t)1 << lg;
    uint32_t m4 = (uint32_t)(bytes / 4 - 1), m8 = (uint32_t)(bytes / 8 - 1), m32 = (uint32_t)(bytes / 32 - 1);
    float a = best_ms([&] { k_gather<0, false><<<grid, 256>>>((const int4*)keys, n4, table, m4, res, sink); });
    float b = best_ms([&] { k_gather<1, false><<<grid, 256>>>((const int4*)keys, n4, table, m8, res, sink); });
    float c = best_ms([&] { k_gather<2, false><<<grid, 256>>>((const int4*)keys, n4, table, m32, res, sink); });
    float aw = best_ms([&] { k_gather<0, true><<<grid, 256>>>((const int4*)keys, n4, table, m4, res, sink); });
    float bw = best_ms([&] { k_gather<1, true><<<grid, 256>>>((const int4*)keys, n4, table, m8, res, sink); });
    float cw = best_ms([&] { k_gather<2, true><<<grid, 256>>>((const int4*)keys, n4, table, m32, res, sink); });
    printf("{\"bench\": \"gather\", \"table_MB\": %.0f, \"g4_ms\": %.4f, \"g8_ms\": %.4f, \"g32_ms\": %.4f, \"g4wb_ms\": %.4f, \"g8wb_ms\": %.4f, \"g32wb_ms\": %.4f, \"g8_Glookups\": %.1f, \"g32wb_Glookups\": %.1f}\n",
           bytes / 1048576.0, a, b, c, aw, bw, cw, n / b / 1e6, n / cw / 1e6);
    fflush(stdout);
  }
  {
    size_t nb = (size_t)1 << 24;               // 16M inserts like C2's build
    for (int lg = 26; lg <= 29; lg++) {
      size_t bytes = (size_t)1 << lg;
      float ms = best_ms([&] { cudaMemsetAsync(table, 0xFF, bytes); k_cas8<<<grid, 256>>>((const int4*)keys, nb / 4, (unsigned long long*)table, (uint32_t)(bytes / 8 - 1), sink); }, 4);
      float msm = best_ms([&] { cudaMemsetAsync(table, 0xFF, bytes); }, 4);
      printf("{\"bench\": \"cas8_16M\", \"table_MB\": %.0f, \"memset+cas_ms\": %.4f, \"memset_ms\": %.4f, \"Ginserts\": %.2f}\n", bytes / 1048576.0, ms, msm, nb / (ms - msm) / 1e6);
    }
  }
  {
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    for (int g : {148, 296}) {
      float ms = best_ms([&] { k_smem<<<g, 512, 131072>>>((const int4*)keys, n4, res, sink); });
      printf("{\"bench\": \"smem_gather8_wb\", \"grid\": %d, \"ms\": %.4f, \"Glookups\": %.1f}\n", g, ms, n / ms / 1e6);
    }
  }
  return 0;
}
