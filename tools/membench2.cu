// membench2.cu — second round of B200 micro-measurements: what bounds a random table lookup?
//   l1      : gather8 from tiny tables (16 KB .. 1 MB): L1TEX-bound or L2-bound?
//   half    : gather8, 32 MB table, on 74 vs 148 CTAs of 1024 threads: SM-side or L2-side limit?
//   g64     : random 64-byte bucket loads (2 x LDG.256) vs table size 64..256 MB, with the 1 GB key stream and 1 GB result stream
//   slice   : P-pass sliced gather8 over a 256 MB table (each pass touches one 256/P MB slice, all keys streamed every pass)
//   tma     : cp.async.bulk 16-byte gathers into shared memory (TMA path) alone and mixed with LSU gathers
//   store   : 1 GB read + 2 GB write with scalar vs vector stores (the write kernel's traffic shape)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol_first())); return r;
}
__device__ __forceinline__ void st_stream(int4* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol_first()));
}
__device__ __forceinline__ void st_stream1(uint32_t* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol_first()));
}
__device__ __forceinline__ uint64_t ld8_keep(const uint64_t* p) {
  uint64_t r; asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol_last())); return r;
}
__device__ __forceinline__ void ld32_keep(const uint64_t* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

__global__ void k_fill_keys(uint32_t* keys, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) keys[i] = mix32((uint32_t)i * 2654435761U + seed);
}

template <bool WB>
__global__ void k_gather8(const int4* __restrict__ keys, size_t n4, const uint64_t* __restrict__ table, uint32_t mask, int4* __restrict__ res, int* sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n4; i += stride) {
    int4 k = ld_stream(keys + i);
    uint64_t v0 = ld8_keep(table + ((uint32_t)k.x & mask)), v1 = ld8_keep(table + ((uint32_t)k.y & mask)), v2 = ld8_keep(table + ((uint32_t)k.z & mask)), v3 = ld8_keep(table + ((uint32_t)k.w & mask));
    if (WB) st_stream(res + i, make_int4((int)v0, (int)v1, (int)v2, (int)v3)); else acc ^= (uint32_t)(v0 ^ v1 ^ v2 ^ v3);
  }
  if (!WB && acc == 0x12345678) *sink = acc;
}

// 64-byte bucket (8 slots): two 32-byte loads, find key among 8 slots, write the row (like the real count kernel would)
__global__ void __launch_bounds__(256) k_gather64(const int4* __restrict__ keys, size_t n4, const uint64_t* __restrict__ table, uint32_t nbuckets_mask, int4* __restrict__ res) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    int4 k = ld_stream(keys + i);
    uint32_t kk[4] = {(uint32_t)k.x, (uint32_t)k.y, (uint32_t)k.z, (uint32_t)k.w};
    uint64_t s[4][8];
    #pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint64_t* b = table + 8 * (size_t)(kk[j] & nbuckets_mask);
      ld32_keep(b, s[j][0], s[j][1], s[j][2], s[j][3]);
      ld32_keep(b + 4, s[j][4], s[j][5], s[j][6], s[j][7]);
    }
    uint32_t r[4];
    #pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t m = 0xFFFFFFFFu;
      #pragma unroll
      for (int e = 0; e < 8; e++) if ((uint32_t)(s[j][e] >> 32) == kk[j]) m = (uint32_t)s[j][e];
      r[j] = m;
    }
    st_stream(res + i, make_int4(r[0], r[1], r[2], r[3]));
  }
}

// sliced: P passes in one launch order; pass p looks up only keys whose slot falls in slice p
__global__ void __launch_bounds__(256) k_slice_pass(const int4* __restrict__ keys, size_t n4, const uint64_t* __restrict__ table, uint32_t mask, uint32_t lo, uint32_t hi,
                                                    unsigned long long* __restrict__ hits) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0, cnt = 0;
  for (; i < n4; i += stride) {
    int4 k = ld_stream(keys + i);
    uint32_t h[4] = {(uint32_t)k.x & mask, (uint32_t)k.y & mask, (uint32_t)k.z & mask, (uint32_t)k.w & mask};
    uint64_t v[4] = {0, 0, 0, 0};
    #pragma unroll
    for (int j = 0; j < 4; j++) if (h[j] >= lo && h[j] < hi) { v[j] = ld8_keep(table + h[j]); cnt++; }
    acc ^= (uint32_t)(v[0] ^ v[1] ^ v[2] ^ v[3]);
  }
  if (acc == 0x12345678) hits[1] = acc;
  atomicAdd(hits, (unsigned long long)cnt);
}

// TMA path: each thread gathers 16 bytes per key with cp.async.bulk into its own smem slot
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int LSU_MIX>   // of every 4 keys, LSU_MIX go through LDG, the rest through TMA
__global__ void __launch_bounds__(256) k_tma_gather(const int4* __restrict__ keys, size_t n4, const uint64_t* __restrict__ table, uint32_t mask2, int4* __restrict__ res) {
  __shared__ alignas(16) uint64_t buf[256 * 4 * 2];
  __shared__ alignas(8) uint64_t bar;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(256)); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t phase = 0;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  size_t iters = (n4 + stride - 1) / stride;
  for (size_t it = 0; it < iters; it++, i += stride) {
    const bool act = i < n4;
    int4 k = act ? ld_stream(keys + i) : make_int4(0, 0, 0, 0);
    uint32_t kk[4] = {(uint32_t)k.x, (uint32_t)k.y, (uint32_t)k.z, (uint32_t)k.w};
    uint64_t lv[4] = {0, 0, 0, 0};
    constexpr int NT = 4 - LSU_MIX;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(16 * NT) : "memory");
    #pragma unroll
    for (int j = 0; j < NT; j++) {
      const uint64_t* src = table + 2 * (size_t)(kk[j] & mask2);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                   :: "r"(smem_u32(&buf[(threadIdx.x * 4 + j) * 2])), "l"(src), "r"(smem_u32(&bar)) : "memory");
    }
    #pragma unroll
    for (int j = NT; j < 4; j++) lv[j] = ld8_keep(table + 2 * (size_t)(kk[j] & mask2));
    // wait
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    }
    phase ^= 1;
    uint32_t r[4];
    #pragma unroll
    for (int j = 0; j < NT; j++) r[j] = (uint32_t)buf[(threadIdx.x * 4 + j) * 2];
    #pragma unroll
    for (int j = NT; j < 4; j++) r[j] = (uint32_t)lv[j];
    if (act) st_stream(res + i, make_int4(r[0], r[1], r[2], r[3]));
    __syncthreads();
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_store(const int4* __restrict__ in, size_t n4, uint32_t* __restrict__ o1, uint32_t* __restrict__ o2) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    int4 v = ld_stream(in + i);
    if (MODE == 0) { st_stream((int4*)o1 + i, v); st_stream((int4*)o2 + i, make_int4(v.x + 1, v.y + 1, v.z + 1, v.w + 1)); }
    else {
      // scalar stores, coalesced per warp: element (4*warp_base + lane + 32*e)
      size_t wbase = (i - (threadIdx.x & 31)) * 4;
      uint32_t lane = threadIdx.x & 31;
      uint32_t x[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
      #pragma unroll
      for (int e = 0; e < 4; e++) { st_stream1(o1 + wbase + lane + 32 * e, x[e]); st_stream1(o2 + wbase + lane + 32 * e, x[e] + 1); }
    }
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

int main() {
  size_t n = (size_t)1 << 28, n4 = n / 4;
  uint32_t* keys; int4* res; uint64_t* table; int* sink; unsigned long long* hits; uint32_t* o2;
  CK(cudaMalloc(&keys, n * 4)); CK(cudaMalloc(&res, n * 4)); CK(cudaMalloc(&o2, n * 4)); CK(cudaMalloc(&table, (size_t)1 << 30)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&hits, 16));
  k_fill_keys<<<148 * 8, 256>>>(keys, n, 12345u); CK(cudaDeviceSynchronize());
  CK(cudaMemset(table, 0xFF, (size_t)1 << 30));
  int grid = 148 * 8;
  for (int lg = 14; lg <= 20; lg++) {
    size_t bytes = (size_t)1 << lg;
    float a = best_ms([&] { k_gather8<false><<<grid, 256>>>((const int4*)keys, n4, table, (uint32_t)(bytes / 8 - 1), res, sink); });
    printf("{\"bench\": \"l1\", \"table_KB\": %zu, \"g8_ms\": %.4f, \"Glookups\": %.1f}\n", bytes / 1024, a, n / a / 1e6);
  }
  for (int g : {37, 74, 148, 296}) {
    float a = best_ms([&] { k_gather8<false><<<g, 1024>>>((const int4*)keys, n4, table, (1u << 22) - 1, res, sink); });
    printf("{\"bench\": \"half\", \"ctas_of_1024\": %d, \"g8_32MB_ms\": %.4f, \"Glookups\": %.1f}\n", g, a, n / a / 1e6);
  }
  for (size_t mb : {32, 64, 80, 96, 112, 128, 144, 160, 192, 256}) {
    // non-power-of-two sizes: use modulo-free mask on the next lower pow2 plus an offset walk — keep it simple: round buckets to pow2 of (mb rounded down), so run pow2 and 1.5x variants via two masks
    size_t bytes = mb << 20;
    size_t nb = bytes / 64;
    // largest pow2 <= nb
    size_t p2 = 1; while (p2 * 2 <= nb) p2 *= 2;
    if (p2 != nb) continue;
    float c = best_ms([&] { k_gather64<<<grid, 256>>>((const int4*)keys, n4, table, (uint32_t)(nb - 1), res); });
    float b8 = best_ms([&] { k_gather8<true><<<grid, 256>>>((const int4*)keys, n4, table, (uint32_t)(bytes / 8 - 1), res, sink); });
    printf("{\"bench\": \"g64\", \"table_MB\": %zu, \"g64wb_ms\": %.4f, \"g8wb_ms\": %.4f}\n", mb, c, b8);
  }
  // sliced passes over a 256 MB table (2^25 slots)
  for (int P : {1, 2, 4, 8}) {
    uint32_t slots = 1u << 25, mask = slots - 1;
    float tot = 0;
    for (int p = 0; p < P; p++) {
      uint32_t lo = (uint32_t)((uint64_t)slots * p / P), hi = (uint32_t)((uint64_t)slots * (p + 1) / P);
      tot += best_ms([&] { k_slice_pass<<<grid, 256>>>((const int4*)keys, n4, table, mask, lo, hi, hits); }, 4);
    }
    printf("{\"bench\": \"slice\", \"P\": %d, \"slice_MB\": %d, \"total_ms\": %.4f}\n", P, 256 / P, tot);
  }
  {
    uint32_t mask2 = (1u << 21) - 1;   // 2^21 x 16 B = 32 MB
    float t0 = best_ms([&] { k_tma_gather<0><<<grid, 256>>>((const int4*)keys, n4, table, mask2, res); });
    float t2 = best_ms([&] { k_tma_gather<2><<<grid, 256>>>((const int4*)keys, n4, table, mask2, res); });
    float t3 = best_ms([&] { k_tma_gather<3><<<grid, 256>>>((const int4*)keys, n4, table, mask2, res); });
    float t4 = best_ms([&] { k_tma_gather<4><<<grid, 256>>>((const int4*)keys, n4, table, mask2, res); });
    printf("{\"bench\": \"tma\", \"table_MB\": 32, \"all_tma_ms\": %.4f, \"half_tma_ms\": %.4f, \"quarter_tma_ms\": %.4f, \"all_lsu_ms\": %.4f}\n", t0, t2, t3, t4);
  }
  {
    float a = best_ms([&] { k_store<0><<<grid, 256>>>((const int4*)keys, n4, (uint32_t*)res, o2); });
    float b = best_ms([&] { k_store<1><<<grid, 256>>>((const int4*)keys, n4, (uint32_t*)res, o2); });
    printf("{\"bench\": \"store_1r2w\", \"vec16_ms\": %.4f, \"scalar4_ms\": %.4f, \"vec_GBps\": %.1f}\n", a, b, 3.0 * n * 4 / a / 1e6);
  }
  return 0;
}
