// hj_common.cuh — shared device helpers for the B200 hash-join kernels (sm_100a only).
//
// Replaces, for the hot path of deveshv-99/mlir-HashJoin:
//   @hash                      join_v1.mlir:206-210   (uint32)key % H  -> multiplicative mix + fast range reduction
//   chained node arrays        join_v1.mlir:25-39     head/lkey/lrow/lnext -> ONE open-addressing slot array
// The result contract (join_v1.mlir:498-500, shared.cpp:139-171) is unchanged: (build_row, probe_row) i32 pairs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "hj_kernels.cuh"

namespace hj {

constexpr uint32_t ROW_NONE = 0xFFFFFFFFu;     // never a valid row id: EMPTY marker lives in the row half of a slot
constexpr uint32_t HJ_MAGIC = 0x424A4832u;     // "2HJB"

// Device-resident table header (first HEADER_BYTES of the table workspace).
struct TableHeader {
  uint32_t magic;
  uint32_t key_bytes;
  unsigned long long n_slots;
  unsigned long long n_rows;
  uint32_t has_dups;        // set by build when two build rows carry the same key
  uint32_t reserved;
};

// ---------------------------------------------------------------------------------------------------------
// hashing
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {          // lowbias32
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z;
}

// Key traits: slot type, packing, hashing.  i32 keys: 8-byte slot  (key << 32 | row).
//                                            i64 keys: 16-byte slot {key, row | 0xFFFFFFFF00000000-free tag}.
template <typename K> struct KeyTraits;

template <> struct KeyTraits<int32_t> {
  using Slot = unsigned long long;
  static constexpr int SLOT_BYTES = 8;
  static constexpr int KEYS_PER_VEC = 4;      // one 16-byte vector load
  // slot index in [0, n_slots): top bits of the mixed key, n_slots <= 2^32
  __device__ __forceinline__ static uint64_t index(int32_t key, uint64_t n_slots) {
    return ((uint64_t)mix32((uint32_t)key) * n_slots) >> 32;
  }
  __device__ __forceinline__ static uint32_t part_hash(int32_t key) { return mix32((uint32_t)key ^ 0x9E3779B9u); }
};

template <> struct KeyTraits<int64_t> {
  using Slot = ulonglong2;
  static constexpr int SLOT_BYTES = 16;
  static constexpr int KEYS_PER_VEC = 2;
  __device__ __forceinline__ static uint64_t index(int64_t key, uint64_t n_slots) {
    return __umul64hi(mix64((uint64_t)key), n_slots);
  }
  __device__ __forceinline__ static uint32_t part_hash(int64_t key) { return (uint32_t)(mix64((uint64_t)key ^ 0x9E3779B97F4A7C15ULL) >> 32); }
};

// ---------------------------------------------------------------------------------------------------------
// cache-hinted memory access (PTX): streams are evict-first and skip L1, the table is evict-last
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_last()  { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }

__device__ __forceinline__ int4 ld_stream_v4(const void* p, uint64_t pol) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_stream_v4(void* p, int4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long ld_table(const unsigned long long* p, uint64_t pol) {
  unsigned long long r;
  asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ ulonglong2 ld_table(const ulonglong2* p, uint64_t pol) {
  ulonglong2 r;
  asm volatile("ld.global.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(pol));
  return r;
}

// 128-bit compare-and-swap for the 16-byte (int64-key) slot: SASS ATOMG.E.CAS.128
__device__ __forceinline__ ulonglong2 atom_cas128(ulonglong2* addr, ulonglong2 cmp, ulonglong2 val) {
  ulonglong2 old;
  asm volatile("{\n .reg .b128 c, v, o;\n mov.b128 c, {%2, %3};\n mov.b128 v, {%4, %5};\n"
               " atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n mov.b128 {%0, %1}, o;\n}"
               : "=l"(old.x), "=l"(old.y) : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr) : "memory");
  return old;
}

// ---------------------------------------------------------------------------------------------------------
// slot helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_slot(int32_t key, uint32_t row) { return ((unsigned long long)(uint32_t)key << 32) | row; }
__device__ __forceinline__ ulonglong2 make_slot(int64_t key, uint32_t row) { return make_ulonglong2((unsigned long long)key, (unsigned long long)row); }
__device__ __forceinline__ bool slot_empty(unsigned long long s) { return (uint32_t)s == ROW_NONE; }
__device__ __forceinline__ bool slot_empty(const ulonglong2& s) { return (uint32_t)s.y == ROW_NONE; }
__device__ __forceinline__ uint32_t slot_row(unsigned long long s) { return (uint32_t)s; }
__device__ __forceinline__ uint32_t slot_row(const ulonglong2& s) { return (uint32_t)s.y; }
__device__ __forceinline__ bool slot_key_eq(unsigned long long s, int32_t key) { return (uint32_t)(s >> 32) == (uint32_t)key; }
__device__ __forceinline__ bool slot_key_eq(const ulonglong2& s, int64_t key) { return s.x == (unsigned long long)key; }

__device__ __forceinline__ unsigned long long slot_cas(unsigned long long* p, unsigned long long v) {
  return atomicCAS(p, ~0ULL, v);
}
__device__ __forceinline__ ulonglong2 slot_cas(ulonglong2* p, ulonglong2 v) {
  return atom_cas128(p, make_ulonglong2(~0ULL, ~0ULL), v);
}

// ---------------------------------------------------------------------------------------------------------
// warp / block scan + reduce (hand-written; blockDim.x multiple of 32, <= 1024)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
  const int lane = threadIdx.x & 31;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) { T t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_reduce_sum(T v) {
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// exclusive scan of one value per thread over the block; returns the thread's exclusive prefix, *total = block sum.
// smem: at least 33 elements of T.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem, T* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  T inc = warp_inclusive_scan(v);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? smem[lane] : T(0);
    T winc = warp_inclusive_scan(w);
    smem[lane] = winc - w;               // exclusive warp base
    if (lane == 31) smem[32] = winc;     // block total
  }
  __syncthreads();
  T base = smem[warp];
  *total = smem[32];
  __syncthreads();                        // smem may be reused by the caller
  return base + inc - v;
}

}  // namespace hj
