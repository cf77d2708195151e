// hj_common.cuh — shared device helpers for the B200 hash-join kernels (sm_100a only).
//
// Replaces, for the hot path of deveshv-99/mlir-HashJoin:
//   @hash                      join_v1.mlir:206-210   (uint32)key % H  -> multiplicative mix + fast range reduction
//   chained node arrays        join_v1.mlir:25-39     head/lkey/lrow/lnext -> ONE bucketised open-addressing table
// The result contract (join_v1.mlir:498-500, shared.cpp:139-171) is unchanged: (build_row, probe_row) i32 pairs.
//
// Table layout (decided by tools/membench*.cu): an SM can issue about one 32-byte L2 sector
// request per clock and DRAM fetches 64 bytes per miss, so the unit of probing is ONE 32-byte bucket (one LDG.256):
// 4 slots of (key:32 | row:32) for i32 keys, 2 slots of (key:64, row:32 + pad) for i64 keys.  Buckets are paired into
// 64-byte groups; an overflowing bucket spills first into its sibling (same DRAM atom), then into the next pair.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "hj_kernels.cuh"

namespace hj {

constexpr uint32_t ROW_NONE = 0xFFFFFFFFu;     // never a valid row id: EMPTY marker lives in the row half of a slot
constexpr uint32_t HJ_MAGIC = 0x424A4833u;
constexpr uint32_t MODE_HASH = 0, MODE_DENSE = 1, MODE_GROUP = 2, MODE_RADIX = 3;
// probe path the last count on a scratch took (scratch counters[CTR_PATH]); the write pass launches exactly that path's kernel
constexpr uint32_t PATH_CACHE = 0, PATH_LISTS = 1, PATH_RANGE = 2, PATH_RADIX = 3;

// Device-resident table header (first HEADER_BYTES of the table workspace). Written by the build kernels, read
// (uniformly) by count/write, so no host round trip is needed to pick the layout.
struct TableHeader {
  uint32_t magic;
  uint32_t key_bytes;
  uint32_t mode;                  // MODE_HASH | MODE_DENSE
  uint32_t has_dups;              // two build rows carry the same key
  uint32_t need_fallback;         // dense build met a duplicate: rebuild as hash
  uint32_t all_present;           // dense, unique and every key of [kmin, kmax] present
  unsigned long long n_pairs;     // hash: number of 64-byte bucket pairs
  unsigned long long n_rows;
  long long kmin, kmax;           // build key range (dense decision)
  unsigned long long dense_range; // kmax - kmin + 1 when dense
  unsigned long long body_bytes;  // capacity behind the header
  unsigned long long pairs_cap;   // pairs the caller's workspace can hold
  unsigned long long rows_offset; // grouped layout: byte offset of the u32 row-id array inside the body
  unsigned long long group_cursor;// grouped layout: bump allocator over the row-id array
  unsigned long long n_groups;    // grouped layout: distinct build keys
  unsigned long long work[4];     // ticket counters of the bounded-grid build kernels (hash, group count, group fill)
  uint32_t policy;                // hjBuildEx flags in force for this table (HJ_POLICY_*): the probe passes follow the TABLE, not a process global
  uint32_t rj_bits1, rj_bits2;    // radix layout: 2^(bits1 + bits2) partitions of the build relation, partition id = top bits of radix_hash(key)
  uint32_t slice_bits;            // inline layout built in table-slice order: the probe relation is partitioned into 2^slice_bits slices first (0 = no)
  unsigned long long rj_keys_off, rj_rows_off, rj_offs_off;   // radix layout: byte offsets (inside the body) of the partitioned keys, row ids, u32 offsets[parts + 1]
  unsigned long long dense_filled;// direct-address build: slots taken after all rows were stored (k_dense_verify); < n_rows proves a duplicate key
};
static_assert(sizeof(TableHeader) <= HEADER_BYTES, "header too large");

// ---------------------------------------------------------------------------------------------------------
// hashing
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {          // lowbias32
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z;
}

template <typename K> struct KeyTraits;

// 32-bit hash of the radix join: the TOP bits pick the partition (two passes of <= 8 bits), the LOW bits the slot of the shared-memory table
template <typename K> __host__ __device__ __forceinline__ uint32_t radix_hash(K key);
template <> __host__ __device__ __forceinline__ uint32_t radix_hash<int32_t>(int32_t key) { return mix32((uint32_t)key + 0x68E31DA4u); }
// i64: one 64-bit multiply folded to 32 bits, one 32-bit multiply-xorshift round (8 instructions; splitmix64 costs 22, and the partition
// kernels hash every tuple four times: ncu showed k_rp_scatter at 2.8 warp instructions per tuple, 31 % issue-bound at 37 % of DRAM)
template <> __host__ __device__ __forceinline__ uint32_t radix_hash<int64_t>(int64_t key) {
  const uint64_t z = (uint64_t)key * 0x9E3779B97F4A7C15ULL;
  uint32_t h = (uint32_t)(z >> 32) ^ (uint32_t)z;
  h *= 0x85EBCA6Bu; h ^= h >> 15; h *= 0xC2B2AE35u;
  return h;
}

template <> struct KeyTraits<int32_t> {
  static constexpr int SLOTS = 4;             // slots per 32-byte bucket
  static constexpr int KEYS_PER_VEC = 4;      // keys per 16-byte vector load
  // (pair index, home half): top bits of the mixed key pick the pair, the low bit picks the half
  __device__ __forceinline__ static void home(int32_t key, uint64_t n_pairs, uint64_t& pair, uint32_t& half) {
    const uint32_t h = mix32((uint32_t)key);
    pair = ((uint64_t)h * n_pairs) >> 32;     // n_pairs <= 2^32
    half = h & 1u;
  }
  __device__ __forceinline__ static uint32_t home_hash32(int32_t key) { return mix32((uint32_t)key); }   // pair index is monotone in this: its top bits name a table slice
  __device__ __forceinline__ static uint32_t part_hash(int32_t key) { return mix32((uint32_t)key ^ 0x9E3779B9u); }
};

template <> struct KeyTraits<int64_t> {
  static constexpr int SLOTS = 2;
  static constexpr int KEYS_PER_VEC = 2;
  __device__ __forceinline__ static void home(int64_t key, uint64_t n_pairs, uint64_t& pair, uint32_t& half) {
    const uint64_t h = mix64((uint64_t)key);
    pair = __umul64hi(h, n_pairs);
    half = (uint32_t)h & 1u;
  }
  __device__ __forceinline__ static uint32_t home_hash32(int64_t key) { return (uint32_t)(mix64((uint64_t)key) >> 32); }
  __device__ __forceinline__ static uint32_t part_hash(int64_t key) { return (uint32_t)(mix64((uint64_t)key ^ 0x9E3779B97F4A7C15ULL) >> 32); }
};

// t-th bucket of a probe sequence: home half, sibling half, then the next pair (home half first again) ...
__device__ __forceinline__ uint64_t probe_bucket(uint64_t pair, uint32_t half, uint32_t t, uint64_t n_pairs) {
  uint64_t p = pair + (t >> 1);
  if (p >= n_pairs) p -= n_pairs;             // t stays far below n_pairs (load factor < 1)
  return 2 * p + ((half ^ t) & 1u);
}

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
__device__ __forceinline__ void st_stream_v2(void* p, uint2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" :: "l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ld_keep_u32(const uint32_t* p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}

// bucket of the radix join's shared-memory table: a multiplicative hash is enough there (the keys of one partition are already a
// pseudo-random subset, picked by radix_hash) and costs 2-4 instructions where mix64 costs ~20 (ncu: k_rj_join ran 55-70 % issue-bound)
template <typename K> __device__ __forceinline__ uint32_t bucket_hash(K key);
template <> __device__ __forceinline__ uint32_t bucket_hash<int32_t>(int32_t key) { return (uint32_t)key * 0x9E3779B1u; }
template <> __device__ __forceinline__ uint32_t bucket_hash<int64_t>(int64_t key) { return (uint32_t)key * 0x9E3779B1u + (uint32_t)((uint64_t)key >> 32) * 0x85EBCA77u; }

// scalar streaming loads (coalesced per warp at any alignment: the partition kernels read segments that start anywhere)
template <typename T> __device__ __forceinline__ T ld_stream(const T* p, uint64_t pol);
template <> __device__ __forceinline__ int32_t ld_stream<int32_t>(const int32_t* p, uint64_t pol) {
  int32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}
template <> __device__ __forceinline__ uint32_t ld_stream<uint32_t>(const uint32_t* p, uint64_t pol) {
  uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}
template <> __device__ __forceinline__ int64_t ld_stream<int64_t>(const int64_t* p, uint64_t pol) {
  long long r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol)); return (int64_t)r;
}

// One 32-byte bucket as four 64-bit words (SASS: LDG.E.ELL2.256).
struct Bucket { unsigned long long w[4]; };
__device__ __forceinline__ Bucket ld_bucket(const void* p) {
  Bucket b;
  asm volatile("ld.global.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.w[0]), "=l"(b.w[1]), "=l"(b.w[2]), "=l"(b.w[3]) : "l"(p));
  return b;
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
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: the probe-key stream and the match-cache stream of the count
// kernel bypass the LSU/L1TEX request path, which is the unit the random table lookups saturate
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :: "r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(smem_addr(smem_src)), "r"(bytes), "l"(pol) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// slot views of a bucket
// ---------------------------------------------------------------------------------------------------------
// i32: slot e = w[e] = key << 32 | row.   i64: slot e = (w[2e] = key, w[2e+1] = row | 0xFFFFFFFF-tag in the low half).
__device__ __forceinline__ bool slot_empty(const Bucket& b, int e, int32_t) { return (uint32_t)b.w[e] == ROW_NONE; }
__device__ __forceinline__ bool slot_empty(const Bucket& b, int e, int64_t) { return (uint32_t)b.w[2 * e + 1] == ROW_NONE; }
__device__ __forceinline__ uint32_t slot_row(const Bucket& b, int e, int32_t) { return (uint32_t)b.w[e]; }
__device__ __forceinline__ uint32_t slot_row(const Bucket& b, int e, int64_t) { return (uint32_t)b.w[2 * e + 1]; }
// true only for an OCCUPIED slot holding `key` (an EMPTY i32 slot has key bits 0xFFFFFFFF, which is a legal key)
__device__ __forceinline__ bool slot_match(const Bucket& b, int e, int32_t key) {
  return (uint32_t)(b.w[e] >> 32) == (uint32_t)key && (uint32_t)b.w[e] != ROW_NONE;
}
__device__ __forceinline__ bool slot_match(const Bucket& b, int e, int64_t key) {
  return b.w[2 * e] == (unsigned long long)key && (uint32_t)b.w[2 * e + 1] != ROW_NONE;
}
// Claim slot e of the bucket at `bp` if it is EMPTY; writes the occupant found (or EMPTY on success) back into b.
__device__ __forceinline__ bool slot_claim(void* bp, Bucket& b, int e, int32_t key, uint32_t row) {
  const unsigned long long mine = ((unsigned long long)(uint32_t)key << 32) | row;
  const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(bp) + e, ~0ULL, mine);
  b.w[e] = old;
  return (uint32_t)old == ROW_NONE;
}
__device__ __forceinline__ bool slot_claim(void* bp, Bucket& b, int e, int64_t key, uint32_t row) {
  const ulonglong2 old = atom_cas128(reinterpret_cast<ulonglong2*>(bp) + e, make_ulonglong2(~0ULL, ~0ULL),
                                     make_ulonglong2((unsigned long long)key, (unsigned long long)row));
  b.w[2 * e] = old.x; b.w[2 * e + 1] = old.y;
  return (uint32_t)old.y == ROW_NONE;
}

// ---------------------------------------------------------------------------------------------------------
// grouped layout (duplicate build keys): 16-byte slots { key:64 ; end_or_cursor:32 | count:32 }, two per bucket, plus
// one u32 row-id array in which the rows of a key are contiguous: rows[end - count, end).  A probe is a unique-key
// probe (stop at the first match or the first non-full bucket) whatever the multiplicity; i32 keys are sign-extended.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_home(long long key, uint64_t n_pairs, uint64_t& pair, uint32_t& half) {
  KeyTraits<int64_t>::home((int64_t)key, n_pairs, pair, half);
}
// payload word of the matching slot (low 32 = end offset, high 32 = count), or 0 when the key is absent
__device__ __forceinline__ unsigned long long group_finish(const char* __restrict__ body, uint64_t n_pairs, long long key, Bucket b) {
  if (b.w[0] == (unsigned long long)key && (uint32_t)b.w[1] != ROW_NONE) return b.w[1];
  if (b.w[2] == (unsigned long long)key && (uint32_t)b.w[3] != ROW_NONE) return b.w[3];
  if ((uint32_t)b.w[3] == ROW_NONE) return 0ULL;                // home bucket not full: the sequence ends here
  uint64_t pair; uint32_t half;
  group_home(key, n_pairs, pair, half);                         // rare: walk on from the recomputed home position
  for (uint32_t t = 1;; t++) {
    b = ld_bucket(body + probe_bucket(pair, half, t, n_pairs) * 32);
    if (b.w[0] == (unsigned long long)key && (uint32_t)b.w[1] != ROW_NONE) return b.w[1];
    if (b.w[2] == (unsigned long long)key && (uint32_t)b.w[3] != ROW_NONE) return b.w[3];
    if ((uint32_t)b.w[3] == ROW_NONE) return 0ULL;
  }
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
// smem: at least 33 elements of T. Ends with a barrier, so smem may be reused immediately.
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
  __syncthreads();
  return base + inc - v;
}
// Ticket scheduling for bounded grids: CTAs take block indices in order from a global counter, exactly like the hardware
// block scheduler would, so the blocks in flight stay one contiguous window of the relation (a plain grid-stride loop lets
// fast CTAs run ahead and the window — and with it the table slice being touched — spreads out).
__device__ __forceinline__ long long next_ticket(unsigned long long* counter, long long* slot) {
  if (threadIdx.x == 0) *slot = (long long)atomicAdd(counter, 1ULL);
  __syncthreads();
  const long long b = *slot;
  __syncthreads();
  return b;
}

// The same with the next ticket fetched AHEAD: thread 0 issues the atomic at the top of an iteration and only needs its result at
// the bottom, so the round trip (~1 us under load) is off the critical path and one barrier per iteration is enough (ncu on the
// i64 build, two barriers and a blocking atomic per 512 keys: barrier stalls 24 warps per issue slot).
//   TicketQueue q (shared);  blk = ticket_first(ctr, &q);  for (it = 0; blk < n; it++) { p = ticket_prefetch(ctr); ...; blk = ticket_advance(&q, it, p); }
struct TicketQueue { long long slot[2]; };
__device__ __forceinline__ long long ticket_first(unsigned long long* counter, TicketQueue* q) {
  if (threadIdx.x == 0) q->slot[0] = (long long)atomicAdd(counter, 1ULL);
  __syncthreads();
  return q->slot[0];
}
__device__ __forceinline__ long long ticket_prefetch(unsigned long long* counter) {
  return threadIdx.x == 0 ? (long long)atomicAdd(counter, 1ULL) : 0LL;
}
__device__ __forceinline__ long long ticket_advance(TicketQueue* q, uint32_t it, long long pending) {
  if (threadIdx.x == 0) q->slot[(it + 1) & 1] = pending;
  __syncthreads();                 // slot[it & 1] was read by everyone before this barrier of the PREVIOUS iteration: safe to reuse next time
  return q->slot[(it + 1) & 1];
}

// Two tickets ahead: the CTA knows its NEXT work unit while it works on the current one, so every thread can pull that unit's input into L2
// (the units are a few tens of KB: without this each one starts with a DRAM round trip nobody overlaps).
//   TicketQueue2 q (shared); ticket2_first(ctr, &q, cur, next); for (it = 0; cur < n; it++) { p = ticket_prefetch(ctr); ...; ticket2_advance(&q, it, p, cur, next); }
struct TicketQueue2 { long long slot[4]; };
__device__ __forceinline__ void ticket2_first(unsigned long long* counter, TicketQueue2* q, long long& cur, long long& next) {
  if (threadIdx.x == 0) { q->slot[0] = (long long)atomicAdd(counter, 1ULL); q->slot[1] = (long long)atomicAdd(counter, 1ULL); }
  __syncthreads();
  cur = q->slot[0]; next = q->slot[1];
}
__device__ __forceinline__ void ticket2_advance(TicketQueue2* q, uint32_t it, long long pending, long long& cur, long long& next) {
  if (threadIdx.x == 0) q->slot[(it + 2) & 3] = pending;
  __syncthreads();                 // slot (it + 2) & 3 was last read two iterations (two barriers) ago
  cur = next; next = q->slot[(it + 2) & 3];
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// every 32-byte sector of [p, p + bytes), spread over the CTA's threads
__device__ __forceinline__ void prefetch_l2_range(const void* p, uint32_t bytes, int threads) {
  const char* c = reinterpret_cast<const char*>(p);
  for (uint32_t o = threadIdx.x * 32u; o < bytes; o += (uint32_t)threads * 32u) prefetch_l2(c + o);
}

template <typename T>
__device__ __forceinline__ T block_reduce_sum(T v, T* smem) {            // result valid in every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_reduce_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T w = lane < nwarps ? smem[lane] : T(0);
  w = warp_reduce_sum(w);
  __syncthreads();
  return w;
}

}  // namespace hj
