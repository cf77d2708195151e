// membench3.cu — can the TMA unit take table lookups off the L1TEX request port?  
//   The LSU path is limited to one 128-byte-line wavefront per clock per SM (membench2 `half`, `gather`), so 2^28 random
//   lookups cost >= 0.92 ms on 148 SMs whatever the table's L2 hit rate.  cp.async.bulk.tensor ... tile::gather4 fetches FOUR
//   arbitrary rows of a 2-D tensor per instruction through the TMA unit.  This bench views a 64 MB u32 table as [2^21 rows][32 B]
//   and measures 2^28 lookups with   g4   : all through gather4
//                                   mix  : per thread and iteration 4 keys through gather4 + 4*(NV-1) through LDG
//                                   lsu  : all through LDG (same loop shape)
//   Every variant streams the keys in and one u32 per key out (the count kernel's traffic shape) and must produce the same checksum.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membench3 tools/membench3.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ int4 ld_stream(const int4* p, uint64_t pol) {
  int4 r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol)); return r;
}
__device__ __forceinline__ void st_stream(int4* p, int4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol));
}
__device__ __forceinline__ uint32_t ld4_keep(const uint32_t* p, uint64_t pol) {
  uint32_t r; asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_fill_keys(uint32_t* keys, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) keys[i] = mix32((uint32_t)i * 2654435761U + seed);
}
__global__ void k_fill_table(uint32_t* t, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) t[i] = mix32((uint32_t)i ^ 0xabcdef01u);
}
__global__ void k_checksum(const uint32_t* r, size_t n, unsigned long long* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (; i < n; i += stride) acc += r[i];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// NV int4 of keys per thread per iteration; the first TV of them are looked up through gather4 (one instruction per int4), the rest through LDG.
// STAGES iterations are in flight per warp (per-warp mbarriers, 32 arrivals each).
template <int NV, int TV, int STAGES, int TMA_THREADS = 256>
__device__ __forceinline__ void lookup_body(const int4* __restrict__ keys, size_t n4, const uint32_t* __restrict__ table, uint32_t mask,
                                            const CUtensorMap& tmap, int4* __restrict__ res, unsigned char* smem_raw) {
  constexpr int WARPS = 8;
  // layout: [STAGES][256 threads][TV][4 rows x 32 B]  (a gather4 destination must be 128-byte aligned)
  int4* buf = reinterpret_cast<int4*>(smem_raw);
  __shared__ alignas(8) uint64_t bar[STAGES > 0 ? STAGES * WARPS : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (TV > 0) {
    if (lane == 0) for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[s * WARPS + warp])), "r"(32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  const uint64_t pf = pol_first(), pl = pol_last();
  const size_t per_iter = (size_t)gridDim.x * 256 * NV;
  const size_t iters = (n4 + per_iter - 1) / per_iter;
  // element index of vector v of iteration it for this thread: coalesced per vector
  auto vec_index = [&](size_t it, int v) { return it * per_iter + ((size_t)blockIdx.x * NV + v) * 256 + threadIdx.x; };

  uint32_t sel[STAGES > 0 ? STAGES : 1][TV > 0 ? TV : 1];      // low 2 bits of the 4 keys of each TMA vector, packed 8 bits each
  uint32_t lres[STAGES > 0 ? STAGES : 1][(NV - TV) > 0 ? (NV - TV) * 4 : 1];

  auto issue = [&](size_t it, int s) {
    int4 k[NV];
    #pragma unroll
    for (int v = 0; v < NV; v++) { const size_t i = vec_index(it, v); k[v] = i < n4 ? ld_stream(keys + i, pf) : make_int4(0, 0, 0, 0); }
    if (TV > 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[s * WARPS + warp])), "r"(128 * TV) : "memory");
    #pragma unroll
    for (int v = 0; v < TV; v++) {
      const uint32_t a = (uint32_t)k[v].x & mask, b = (uint32_t)k[v].y & mask, c = (uint32_t)k[v].z & mask, d = (uint32_t)k[v].w & mask;
      sel[s][v] = (a & 7) | ((b & 7) << 8) | ((c & 7) << 16) | ((d & 7) << 24);
      int4* dst = buf + (((size_t)s * TMA_THREADS + threadIdx.x) * TV + v) * 8;
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   :: "r"(smem_u32(dst)), "l"(&tmap), "r"(smem_u32(&bar[s * WARPS + warp])), "r"(0), "r"((int)(a >> 3)), "r"((int)(b >> 3)), "r"((int)(c >> 3)), "r"((int)(d >> 3)) : "memory");
    }
    #pragma unroll
    for (int v = TV; v < NV; v++) {
      lres[s][(v - TV) * 4 + 0] = ld4_keep(table + ((uint32_t)k[v].x & mask), pl);
      lres[s][(v - TV) * 4 + 1] = ld4_keep(table + ((uint32_t)k[v].y & mask), pl);
      lres[s][(v - TV) * 4 + 2] = ld4_keep(table + ((uint32_t)k[v].z & mask), pl);
      lres[s][(v - TV) * 4 + 3] = ld4_keep(table + ((uint32_t)k[v].w & mask), pl);
    }
  };
  auto retire = [&](size_t it, int s, uint32_t parity) {
    if (TV > 0) {
      uint32_t ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[s * WARPS + warp])), "r"(parity) : "memory");
    }
    #pragma unroll
    for (int v = 0; v < NV; v++) {
      const size_t i = vec_index(it, v);
      int4 r;
      if (v < TV) {
        const uint32_t* rows = reinterpret_cast<const uint32_t*>(buf + (((size_t)s * TMA_THREADS + threadIdx.x) * TV + v) * 8);
        const uint32_t sl = sel[s][v];
        r = make_int4(rows[0 + (sl & 7)], rows[8 + ((sl >> 8) & 7)], rows[16 + ((sl >> 16) & 7)], rows[24 + ((sl >> 24) & 7)]);
      } else {
        r = make_int4(lres[s][(v - TV) * 4 + 0], lres[s][(v - TV) * 4 + 1], lres[s][(v - TV) * 4 + 2], lres[s][(v - TV) * 4 + 3]);
      }
      if (i < n4) st_stream(res + i, r, pf);
    }
    __syncwarp();   // everybody has read its smem slot before the slot's next TMA write is issued
  };

  if constexpr (STAGES <= 1) {
    for (size_t it = 0; it < iters; it++) { issue(it, 0); retire(it, 0, it & 1); }
  } else {
    // software pipeline: STAGES iterations in flight (static stage indices through full unrolling of the stage loop)
    #pragma unroll
    for (int s = 0; s < STAGES - 1; s++) if ((size_t)s < iters) issue(s, s);
    for (size_t base = 0; base < iters; base += STAGES) {
      #pragma unroll
      for (int s = 0; s < STAGES; s++) {
        const size_t it = base + s;
        if (it < iters) {
          if (it + STAGES - 1 < iters) issue(it + STAGES - 1, (s + STAGES - 1) % STAGES);
          retire(it, s, (uint32_t)((it / STAGES) & 1));
        }
      }
    }
  }
}

template <int NV, int TV, int STAGES>
__global__ void __launch_bounds__(256) k_lookup(const int4* __restrict__ keys, size_t n4, const uint32_t* __restrict__ table, uint32_t mask,
                                                const __grid_constant__ CUtensorMap tmap, int4* __restrict__ res) {
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  lookup_body<NV, TV, STAGES>(keys, n4, table, mask, tmap, res, smem_dyn);
}

// warp-specialised: warps [0, TW) look their keys up through gather4 only, the others through LDG only (same keys per warp)
template <int NV, int TW, int STAGES>
__global__ void __launch_bounds__(256) k_lookup_ws(const int4* __restrict__ keys, size_t n4, const uint32_t* __restrict__ table, uint32_t mask,
                                                   const __grid_constant__ CUtensorMap tmap, int4* __restrict__ res) {
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  if ((threadIdx.x >> 5) < TW) lookup_body<NV, NV, STAGES, TW * 32>(keys, n4, table, mask, tmap, res, smem_dyn);
  else lookup_body<NV, 0, STAGES>(keys, n4, table, mask, tmap, res, smem_dyn);
}

template <typename F>
static float best_ms(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; r++) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r > 0) best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NV, int TV, int STAGES>
static void run(const char* name, int ctas_per_sm, const uint32_t* keys, size_t n, const uint32_t* table, uint32_t mask, const CUtensorMap& tmap, uint32_t* res, unsigned long long* dsum) {
  const size_t smem = (size_t)(STAGES > 0 ? STAGES : 1) * 256 * (TV > 0 ? TV : 0) * 128;
  CK(cudaFuncSetAttribute(k_lookup<NV, TV, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lookup<NV, TV, STAGES>, 256, smem));
  const int per_sm = std::min(occ, ctas_per_sm);
  const int grid = 148 * per_sm;
  CK(cudaMemset(res, 0, n * 4));
  float t = best_ms([&] { k_lookup<NV, TV, STAGES><<<grid, 256, smem>>>((const int4*)keys, n / 4, table, mask, tmap, (int4*)res); });
  CK(cudaMemset(dsum, 0, 8));
  k_checksum<<<1184, 256>>>(res, n, dsum);
  unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
  printf("{\"bench\": \"%s\", \"keys_per_iter\": %d, \"tma_keys_per_iter\": %d, \"stages\": %d, \"ctas_per_sm\": %d, \"occ_limit\": %d, \"ms\": %.4f, \"G_lookups_per_s\": %.1f, \"checksum\": %llu}\n",
         name, NV * 4, TV * 4, STAGES, per_sm, occ, t, n / t / 1e6, h);
  fflush(stdout);
}

template <int NV, int TW, int STAGES>
static void run_ws(const char* name, const uint32_t* keys, size_t n, const uint32_t* table, uint32_t mask, const CUtensorMap& tmap, uint32_t* res, unsigned long long* dsum) {
  const size_t smem = (size_t)STAGES * TW * 32 * NV * 128;
  CK(cudaFuncSetAttribute(k_lookup_ws<NV, TW, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lookup_ws<NV, TW, STAGES>, 256, smem));
  const int per_sm = std::min(occ, 8);
  CK(cudaMemset(res, 0, n * 4));
  float t = best_ms([&] { k_lookup_ws<NV, TW, STAGES><<<148 * per_sm, 256, smem>>>((const int4*)keys, n / 4, table, mask, tmap, (int4*)res); });
  CK(cudaMemset(dsum, 0, 8));
  k_checksum<<<1184, 256>>>(res, n, dsum);
  unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
  printf("{\"bench\": \"%s\", \"keys_per_iter\": %d, \"tma_warps\": %d, \"stages\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"G_lookups_per_s\": %.1f, \"checksum\": %llu}\n",
         name, NV * 4, TW, STAGES, per_sm, t, n / t / 1e6, h);
  fflush(stdout);
}

// Streams through the TMA unit, lookups through LDG: persistent CTAs, tile = 2048 keys (8 KB in, 8 KB out).
//   LT: the tile's keys arrive by ONE cp.async.bulk (global -> shared, mbarrier), double-buffered
//   ST: the tile's results leave by ONE cp.async.bulk (shared -> global, bulk_group), triple-buffered; one block barrier per tile
template <bool LT, bool ST>
__global__ void __launch_bounds__(256) k_lookup_bulk(const int4* __restrict__ keys, size_t n4, const uint32_t* __restrict__ table, uint32_t mask, int4* __restrict__ res) {
  __shared__ __align__(128) int4 kbuf[2][512];
  __shared__ __align__(128) int4 rbuf[3][512];
  __shared__ __align__(8) uint64_t bar[2];
  const size_t ntiles = n4 / 512;                     // bench sizes are multiples of the tile
  if (threadIdx.x == 0) { for (int s = 0; s < 2; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[s])), "r"(1)); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const uint64_t pf = pol_first(), pl = pol_last();
  auto load_tile = [&](size_t tile, int b) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[b])), "r"(8192) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], 8192, [%2], %3;"
                 :: "r"(smem_u32(&kbuf[b][0])), "l"(keys + tile * 512), "r"(smem_u32(&bar[b])), "l"(pf) : "memory");
  };
  size_t tile = blockIdx.x;
  if (LT && threadIdx.x == 0 && tile < ntiles) load_tile(tile, 0);
  uint32_t it = 0;
  for (; tile < ntiles; tile += gridDim.x, it++) {
    const int kb = it & 1, rb = it % 3;
    int4 k0, k1;
    if (LT) {
      if (threadIdx.x == 0 && tile + gridDim.x < ntiles) load_tile(tile + gridDim.x, kb ^ 1);
      uint32_t ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[kb])), "r"((it >> 1) & 1) : "memory");
      k0 = kbuf[kb][threadIdx.x]; k1 = kbuf[kb][256 + threadIdx.x];
    } else {
      k0 = ld_stream(keys + tile * 512 + threadIdx.x, pf); k1 = ld_stream(keys + tile * 512 + 256 + threadIdx.x, pf);
    }
    int4 r0, r1;
    r0.x = ld4_keep(table + ((uint32_t)k0.x & mask), pl); r0.y = ld4_keep(table + ((uint32_t)k0.y & mask), pl);
    r0.z = ld4_keep(table + ((uint32_t)k0.z & mask), pl); r0.w = ld4_keep(table + ((uint32_t)k0.w & mask), pl);
    r1.x = ld4_keep(table + ((uint32_t)k1.x & mask), pl); r1.y = ld4_keep(table + ((uint32_t)k1.y & mask), pl);
    r1.z = ld4_keep(table + ((uint32_t)k1.z & mask), pl); r1.w = ld4_keep(table + ((uint32_t)k1.w & mask), pl);
    if (ST) {
      rbuf[rb][threadIdx.x] = r0; rbuf[rb][256 + threadIdx.x] = r1;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], 8192, %2;" :: "l"(res + tile * 512), "r"(smem_u32(&rbuf[rb][0])), "l"(pf) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
    } else {
      st_stream(res + tile * 512 + threadIdx.x, r0, pf); st_stream(res + tile * 512 + 256 + threadIdx.x, r1, pf);
      if (LT) __syncthreads();          // the key buffer is reloaded two tiles later
    }
  }
  if (ST && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <bool LT, bool ST>
static void run_bulk(const char* name, int per_sm, const uint32_t* keys, size_t n, const uint32_t* table, uint32_t mask, uint32_t* res, unsigned long long* dsum) {
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lookup_bulk<LT, ST>, 256, 0));
  per_sm = std::min(per_sm, occ);
  CK(cudaMemset(res, 0, n * 4));
  float t = best_ms([&] { k_lookup_bulk<LT, ST><<<148 * per_sm, 256>>>((const int4*)keys, n / 4, table, mask, (int4*)res); });
  CK(cudaMemset(dsum, 0, 8));
  k_checksum<<<1184, 256>>>(res, n, dsum);
  unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
  printf("{\"bench\": \"%s\", \"loads_tma\": %d, \"stores_tma\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"G_lookups_per_s\": %.1f, \"checksum\": %llu}\n", name, (int)LT, (int)ST, per_sm, t, n / t / 1e6, h);
  fflush(stdout);
}

int main(int argc, char** argv) {
  size_t n = (size_t)1 << 28;
  if (argc > 1) n = (size_t)1 << atoi(argv[1]);
  const size_t tn = (size_t)1 << 24;                // 64 MB of u32
  uint32_t *keys, *table, *res; unsigned long long* dsum;
  CK(cudaMalloc(&keys, n * 4)); CK(cudaMalloc(&table, tn * 4)); CK(cudaMalloc(&res, n * 4)); CK(cudaMalloc(&dsum, 8));
  k_fill_keys<<<1184, 256>>>(keys, n, 7u);
  k_fill_table<<<1184, 256>>>(table, tn);
  CK(cudaDeviceSynchronize());

  EncodeTiled enc = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  if (!enc) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {8, tn / 8}; cuuint64_t gstride[1] = {32}; cuuint32_t box[2] = {8, 1}; cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, table, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  const uint32_t mask = (uint32_t)tn - 1;

  for (int per_sm : {4, 5, 8}) {
    run_bulk<false, false>("bulk_none", per_sm, keys, n, table, mask, res, dsum);
    run_bulk<true, false>("bulk_loads", per_sm, keys, n, table, mask, res, dsum);
    run_bulk<false, true>("bulk_stores", per_sm, keys, n, table, mask, res, dsum);
    run_bulk<true, true>("bulk_both", per_sm, keys, n, table, mask, res, dsum);
  }
  if (argc > 2) return 0;
  run_ws<2, 2, 1>("ws_2of8_nv2_s1", keys, n, table, mask, tmap, res, dsum);
  run_ws<2, 3, 1>("ws_3of8_nv2_s1", keys, n, table, mask, tmap, res, dsum);
  run_ws<2, 4, 1>("ws_4of8_nv2_s1", keys, n, table, mask, tmap, res, dsum);
  run_ws<1, 2, 2>("ws_2of8_nv1_s2", keys, n, table, mask, tmap, res, dsum);
  run_ws<1, 3, 2>("ws_3of8_nv1_s2", keys, n, table, mask, tmap, res, dsum);
  run_ws<1, 4, 2>("ws_4of8_nv1_s2", keys, n, table, mask, tmap, res, dsum);
  run_ws<2, 3, 2>("ws_3of8_nv2_s2", keys, n, table, mask, tmap, res, dsum);
  run_ws<1, 3, 1>("ws_3of8_nv1_s1", keys, n, table, mask, tmap, res, dsum);
  run_ws<1, 8, 1>("ws_8of8_nv1_s1", keys, n, table, mask, tmap, res, dsum);
  run<2, 0, 0>("lsu",  8, keys, n, table, mask, tmap, res, dsum);
  run<2, 0, 2>("lsu_pipe2", 8, keys, n, table, mask, tmap, res, dsum);
  run<1, 1, 1>("g4_s1", 8, keys, n, table, mask, tmap, res, dsum);
  run<1, 1, 2>("g4_s2", 8, keys, n, table, mask, tmap, res, dsum);
  run<1, 1, 4>("g4_s4", 8, keys, n, table, mask, tmap, res, dsum);
  run<2, 2, 2>("g4x2_s2", 8, keys, n, table, mask, tmap, res, dsum);
  run<2, 1, 2>("mix_4t_4l_s2", 8, keys, n, table, mask, tmap, res, dsum);
  run<2, 1, 4>("mix_4t_4l_s4", 8, keys, n, table, mask, tmap, res, dsum);
  run<3, 1, 2>("mix_4t_8l_s2", 8, keys, n, table, mask, tmap, res, dsum);
  run<3, 1, 4>("mix_4t_8l_s4", 8, keys, n, table, mask, tmap, res, dsum);
  run<4, 1, 2>("mix_4t_12l_s2", 8, keys, n, table, mask, tmap, res, dsum);
  run<5, 1, 2>("mix_4t_16l_s2", 8, keys, n, table, mask, tmap, res, dsum);
  return 0;
}
