// membench5.cu — what does ranking one partition tile cost?  A CTA of 1 024 threads ranks 8 192 tuples by an 8-bit digit (position of
// each tuple in digit order) and stages the keys in shared memory in that order — the first half of k_rp_scatter (hj_partition.cu).
// Variants of the rank step:
//   0  one returning shared-memory atomic per tuple on a CTA-wide counter array (what k_rp_scatter does)
//   1  ballot multisplit (8 ballots per tuple) + warp-private counters touched by one leader lane per digit with plain loads / stores
//   2  __match_any_sync instead of the 8 ballots, same counters
//   3  non-returning atomics only (a histogram: no ranks, calibration)
//   4  no ranking at all, keys staged at their own index (floor of everything else in the loop)
//   5  returning atomics on warp-private counters (is the CTA-wide contention the cost?)
// Prints cycles per tile per SM (one CTA per SM) at the clock the run saw.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membench5 tools/membench5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int THREADS = 1024, ITEMS = 8, TILE = THREADS * ITEMS, FAN = 256, WARPS = THREADS / 32;

struct Smem {
  unsigned long long skeys[TILE];
  uint32_t cnt[FAN], lbase[FAN];
  uint32_t wc[WARPS][FAN];       // warp-private counters (variants 1, 2, 5)
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v) {
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v += t; }
  return v;
}

template <int V>
__global__ void __launch_bounds__(THREADS, 1) k_rank(int tiles, unsigned long long* sink, uint32_t* check) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = (1u << lane) - 1;
  for (int d = threadIdx.x; d < FAN; d += THREADS) sm.cnt[d] = 0;
  for (int i = threadIdx.x; i < WARPS * FAN; i += THREADS) (&sm.wc[0][0])[i] = 0;
  __syncthreads();
  unsigned long long acc = 0; uint32_t bad = 0;
  for (int t = 0; t < tiles; t++) {
    uint32_t dg[ITEMS], pr[ITEMS]; unsigned long long key[ITEMS];
    #pragma unroll
    for (int e = 0; e < ITEMS; e++) {
      const uint32_t h = mix32((blockIdx.x * tiles + t) * TILE + e * THREADS + threadIdx.x);
      key[e] = ((unsigned long long)h << 32) | (e * THREADS + threadIdx.x); dg[e] = h >> 24;
    }
    if (V == 0) {
      #pragma unroll
      for (int e = 0; e < ITEMS; e++) pr[e] = atomicAdd(&sm.cnt[dg[e]], 1u);
    } else if (V == 3) {
      #pragma unroll
      for (int e = 0; e < ITEMS; e++) { atomicAdd(&sm.cnt[dg[e]], 1u); pr[e] = 0; }
    } else if (V == 5) {
      #pragma unroll
      for (int e = 0; e < ITEMS; e++) pr[e] = atomicAdd(&sm.wc[warp][dg[e]], 1u);
    } else if (V == 1 || V == 2) {
      #pragma unroll
      for (int e = 0; e < ITEMS; e++) {
        uint32_t mask;
        if (V == 2) mask = __match_any_sync(0xffffffffu, dg[e]);
        else {
          mask = 0xffffffffu;
          #pragma unroll
          for (int b = 0; b < 8; b++) { const uint32_t bal = __ballot_sync(0xffffffffu, (dg[e] >> b) & 1u); mask &= ((dg[e] >> b) & 1u) ? bal : ~bal; }
        }
        const int leader = __ffs(mask) - 1;
        uint32_t old = 0;
        if ((int)lane == leader) { old = sm.wc[warp][dg[e]]; sm.wc[warp][dg[e]] = old + __popc(mask); }
        old = __shfl_sync(0xffffffffu, old, leader);
        pr[e] = old + __popc(mask & lt);
        __syncwarp();
      }
    } else {
      #pragma unroll
      for (int e = 0; e < ITEMS; e++) pr[e] = 0;
    }
    __syncthreads();
    if (V == 1 || V == 2 || V == 5) {                       // per digit: exclusive scan over the warps' counters, total into cnt
      if (threadIdx.x < FAN) {
        uint32_t run = 0;
        #pragma unroll 8
        for (int w = 0; w < WARPS; w++) { const uint32_t v = sm.wc[w][threadIdx.x]; sm.wc[w][threadIdx.x] = run; run += v; }
        sm.cnt[threadIdx.x] = run;
      }
      __syncthreads();
    }
    if (threadIdx.x < 32) {
      uint32_t v[FAN / 32], sum = 0;
      #pragma unroll
      for (int q = 0; q < FAN / 32; q++) { v[q] = sm.cnt[threadIdx.x * (FAN / 32) + q]; sum += v[q]; }
      uint32_t run = warp_scan_incl(sum) - sum;
      #pragma unroll
      for (int q = 0; q < FAN / 32; q++) { const int d = threadIdx.x * (FAN / 32) + q; sm.lbase[d] = run; sm.cnt[d] = 0; run += v[q]; }
    }
    __syncthreads();
    #pragma unroll
    for (int e = 0; e < ITEMS; e++) {
      uint32_t pos;
      if (V == 0) pos = sm.lbase[dg[e]] + pr[e];
      else if (V == 1 || V == 2 || V == 5) pos = sm.lbase[dg[e]] + sm.wc[warp][dg[e]] + pr[e];
      else pos = e * THREADS + threadIdx.x;
      sm.skeys[pos] = key[e];
    }
    __syncthreads();
    if (V == 1 || V == 2 || V == 5) {
      uint4* z = reinterpret_cast<uint4*>(&sm.wc[0][0]);
      for (int i = threadIdx.x; i < WARPS * FAN / 4; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    #pragma unroll
    for (int e = 0; e < ITEMS; e++) {                       // read back in order: digits must be non-decreasing, every slot written
      const uint32_t i = e * THREADS + threadIdx.x;
      const unsigned long long k = sm.skeys[i];
      acc += k;
      if (V != 3 && V != 4 && i > 0 && (uint32_t)(sm.skeys[i - 1] >> 56) > (uint32_t)(k >> 56)) bad++;
    }
    __syncthreads();
  }
  if (acc == 0x1234567ULL) *sink = acc;
  if (bad) atomicAdd(check, bad);
}

template <int V>
static void run(const char* name, int tiles, unsigned long long* sink, uint32_t* check, double mhz) {
  auto kern = k_rank<V>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  CK(cudaMemset(check, 0, 4));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a));
    kern<<<148, THREADS, sizeof(Smem)>>>(tiles, sink, check);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  uint32_t bad; CK(cudaMemcpy(&bad, check, 4, cudaMemcpyDeviceToHost));
  const double cyc = best * 1e-3 * mhz * 1e6 / tiles;
  printf("{\"variant\": \"%s\", \"ms\": %.3f, \"tiles_per_cta\": %d, \"cycles_per_tile\": %.0f, \"cycles_per_tuple\": %.3f, \"ms_per_2^28_tuples_148sm\": %.3f, \"order_violations\": %u}\n",
         name, best, tiles, cyc, cyc / TILE, best / tiles * (268435456.0 / TILE / 148.0), bad);
}

int main() {
  int mhz_k = 0; CK(cudaDeviceGetAttribute(&mhz_k, cudaDevAttrClockRate, 0));
  const double mhz = mhz_k / 1000.0;
  unsigned long long* sink; uint32_t* check;
  CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&check, 4));
  const int tiles = 400;
  run<4>("stage only", tiles, sink, check, mhz);
  run<3>("atomics without return (histogram)", tiles, sink, check, mhz);
  run<0>("returning atomics, CTA-wide counters", tiles, sink, check, mhz);
  run<5>("returning atomics, warp-private counters", tiles, sink, check, mhz);
  run<1>("ballot multisplit, warp-private counters", tiles, sink, check, mhz);
  run<2>("match_any, warp-private counters", tiles, sink, check, mhz);
  return 0;
}
