// hj_kernels.cu — hand-written sm_100a kernels for the hash-join hot path.
//
//   K0 clear   + K1 build   replace  @initializeHT / @build + @insertNodeInHashTable   join_v1.mlir:180-202, 213-277
//   K2 count   + K3 scan    replace  @count (chain walk + thread-0 serial scan + i32 atomic)  join_v1.mlir:288-425
//   K4 write                replaces @probe (second chain walk, scattered 4-byte stores)   join_v1.mlir:436-521,
//                                    and join_v2's shared-memory staging                   join_v2.mlir:450-604
//   K5 radix partition      new (the reference lists partitioned joins as left out, projectDescription.md:24)
//   K6 digest / generators  verification + seeded inputs (the reference's are unseeded, shared.cpp:62,86-87)
//
// Data structure: one open-addressing table, linear probing, EMPTY = row half 0xFFFFFFFF, insertion by a single
// packed CAS (8-byte slot for i32 keys, 16-byte slot + ATOMG.CAS.128 for i64 keys).  Output contract is the
// reference's: two i32 columns (build_row, probe_row), any order (join_v1.mlir:498-500, shared.cpp:168-171).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "hj_common.cuh"
#include "hj_kernels.cuh"

namespace hj {

// =========================================================================================================
// geometry
// =========================================================================================================
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// 64-byte bucket pairs for n_rows build rows: load factor 0.5 (probe sequences: 1.05 buckets on average, 99th percentile 2;
// at 0.8 the 99th percentile is 8 and a warp waits for its slowest lane). Footprint beyond L2 is handled by probing in
// table-slice order (LOCALITY_* below), not by packing the table tighter.
int64_t preferred_pairs(int64_t n_rows, int key_bytes) {
  const int64_t slots_per_pair = key_bytes == 4 ? 8 : 4;
  const int64_t pairs = (2 * n_rows + slots_per_pair - 1) / slots_per_pair;
  return pairs < 4 ? 4 : pairs;
}
// The workspace must also hold the grouped layout a build with duplicate keys is rebuilt into: a u32 row-id array plus
// 16-byte slots at load factor <= 0.8 for up to n_rows distinct keys.
static int64_t table_part_bytes(int64_t n_rows, int key_bytes) {
  const int64_t inline_bytes = preferred_pairs(n_rows, key_bytes) * 64;
  const int64_t group_bytes = round_up(n_rows * 4, 64) + ((5 * n_rows / 4 + 3) / 4 + 4) * 64;
  return round_up(std::max(inline_bytes, group_bytes), 256);
}
// Tables beyond L2 reach (tools/membench: random lookups stop hitting L2 past ~64 MB of table) are not hash tables in global memory at
// all: both relations are radix-partitioned (K5, two passes) until a build partition fits a shared-memory table, and joined partition
// by partition (K7, hj_radix.cu). The workspace still holds room for the bucketised layouts (policy bit HJ_POLICY_RADIX off).
constexpr int64_t LOCALITY_MIN_BYTES = (int64_t)48 << 20;
bool table_is_big(int64_t n_rows, int key_bytes) { return preferred_pairs(n_rows, key_bytes) * 64 > LOCALITY_MIN_BYTES; }
// Between "fits L2" and "too big to slice": an inline table of at most 1 GB is still built and probed as ONE hash table in global memory,
// but in TABLE-SLICE order — both relations are partitioned once (K5, 16 .. 256 slices of <= 8 MB of table) on the hash bits that pick the
// bucket pair (of the grouped table for duplicate keys), so the CTAs in flight touch a few L2-resident slices at a time. (2 MB slices measured no better in the probe kernel — 1.58 ms
// for 2^28 lookups either way — and cost the partition pass its long runs: 1.23 ms at 128 slices.) One partition pass per side instead of the radix layout's
// two; beyond 1 GB the slices themselves outgrow L2 (round 1: 33 MB slices, 50 % hits) and the radix layout takes over.
constexpr int64_t SLICED_MAX_BYTES = (int64_t)1 << 30, SLICE_BYTES = (int64_t)8 << 20;
static bool table_sliceable(int64_t n_rows, int key_bytes) { return table_is_big(n_rows, key_bytes) && preferred_pairs(n_rows, key_bytes) * 64 <= SLICED_MAX_BYTES; }
static int slice_bits_for(int64_t n_rows, int key_bytes) {
  int bits = 4;
  while (bits < 8 && (preferred_pairs(n_rows, key_bytes) * 64 >> bits) > SLICE_BYTES) bits++;
  return bits;
}
// slice-ordered build: a partitioned copy of the build relation behind the table body: [keys][row ids][offsets u32 x 257][partition workspace]
static int64_t sliced_extra_bytes(int64_t n, int key_bytes) { return round_up(n * key_bytes, 256) + round_up(n * 4, 256) + round_up(257 * 4, 256) + round_up(slice_partition_workspace_bytes(n, 8), 256); }
int64_t table_bytes(int64_t n_rows, int key_bytes) {
  int64_t body = table_part_bytes(n_rows, key_bytes);
  if (table_sliceable(n_rows, key_bytes)) body += sliced_extra_bytes(n_rows, key_bytes);
  if (table_is_big(n_rows, key_bytes)) body = std::max(body, round_up(radix_table_bytes(n_rows, key_bytes), 256));
  return HEADER_BYTES + body;
}
int64_t num_chunks(int64_t n_probe, int key_bytes) {
  const int64_t c = chunk_keys(key_bytes);
  return (n_probe + c - 1) / c;
}
constexpr int SCRATCH_SCAN_BLOCKS = 264;     // one total per 16 384 entries for the two-level scan
// entries of the offsets array: one per chunk (bucketised / direct-address layouts) or one per work item (radix layout), whichever is more
static int64_t scratch_offsets_len(int64_t n_probe, int key_bytes) { return std::max(num_chunks(n_probe, key_bytes), radix_max_items(n_probe)); }
static int64_t scratch_core_bytes(int64_t n_probe, int key_bytes) {
  const int64_t nc = num_chunks(n_probe, key_bytes);
  return 2 * round_up(nc * chunk_keys(key_bytes) * 4, 256) + round_up((scratch_offsets_len(n_probe, key_bytes) + 1 + SCRATCH_COUNTERS + SCRATCH_SCAN_BLOCKS) * 8, 256) + round_up(nc * (BLOCK_THREADS / 32) * 4, 256);
}
// match cache + hit positions + offsets + counters + the radix layout's partitioned copy of the probe relation
int64_t scratch_bytes(int64_t n_probe, int key_bytes) { return scratch_core_bytes(n_probe, key_bytes) + radix_scratch_bytes(n_probe, key_bytes) + 256; }
ScratchView scratch_view(void* scratch, int64_t n_probe, int key_bytes) {
  ScratchView v;
  char* base = reinterpret_cast<char*>(scratch);
  v.nchunks = num_chunks(n_probe, key_bytes);
  v.mcache = reinterpret_cast<uint32_t*>(base);
  const int64_t cache_bytes = round_up(v.nchunks * chunk_keys(key_bytes) * 4, 256);
  v.hit_list = reinterpret_cast<uint2*>(base);                  // the same bytes as the match cache plus as much again behind it
  v.run_start = reinterpret_cast<uint32_t*>(base + cache_bytes); // grouped layout: first row-id slot of each probe row's run (second half)
  v.chunk_offsets = reinterpret_cast<unsigned long long*>(base + 2 * cache_bytes);
  const int64_t noff = scratch_offsets_len(n_probe, key_bytes);
  v.counters = v.chunk_offsets + noff + 1;                      // CTR_* below
  v.scan_sums = v.counters + SCRATCH_COUNTERS;
  v.warp_counts = reinterpret_cast<uint32_t*>(base + 2 * cache_bytes + round_up((noff + 1 + SCRATCH_COUNTERS + SCRATCH_SCAN_BLOCKS) * 8, 256));
  v.radix = base + scratch_core_bytes(n_probe, key_bytes);
  return v;
}

// =========================================================================================================
// K0/K1  build:  init header -> min/max of the build keys -> decide layout -> clear -> dense build
//                -> (duplicate in dense mode: switch to hash, clear again) -> hash build.
// Every decision is taken on the device and read back by later kernels through the header: no host round trip.
// =========================================================================================================
constexpr int DENSE_MAX_FACTOR = 4;     // direct addressing when (kmax - kmin + 1) <= 4 x build rows and it fits

__global__ void k_init_header(TableHeader* hdr, uint32_t key_bytes, unsigned long long n_rows, unsigned long long body_bytes, unsigned long long pairs, uint32_t policy) {
  hdr->policy = policy; hdr->slice_bits = 0; hdr->rj_bits1 = hdr->rj_bits2 = 0; hdr->rj_keys_off = hdr->rj_rows_off = hdr->rj_offs_off = 0;
  hdr->magic = HJ_MAGIC; hdr->key_bytes = key_bytes; hdr->mode = MODE_HASH; hdr->has_dups = 0; hdr->need_fallback = 0; hdr->all_present = 0;
  hdr->n_pairs = pairs; hdr->n_rows = n_rows; hdr->kmin = 0x7FFFFFFFFFFFFFFFLL; hdr->kmax = -0x7FFFFFFFFFFFFFFFLL - 1;
  hdr->dense_range = 0; hdr->body_bytes = body_bytes; hdr->pairs_cap = body_bytes / 64;
  hdr->rows_offset = 0; hdr->group_cursor = 0; hdr->n_groups = 0;
  hdr->work[0] = hdr->work[1] = hdr->work[2] = hdr->work[3] = 0; hdr->dense_filled = 0;
}

template <typename K>
__global__ void __launch_bounds__(BLOCK_THREADS) k_minmax(const K* __restrict__ R, int64_t nR, TableHeader* hdr) {
  __shared__ long long smin[32], smax[32];
  {
    // k_sample_dups ran first and left the min / max of its 65 536 sampled keys in the header: bounds of a SUBSET, so a sampled range
    // already too wide for the direct-address layout proves the full range is too (exact, no false negatives) and the scan of the whole
    // column is skipped (2^28 i64 keys: 0.30 ms)
    const long long slo = hdr->kmin, shi = hdr->kmax;
    const unsigned long long srange = (unsigned long long)shi - (unsigned long long)slo + 1ULL;
    if (slo <= shi && (srange == 0 || srange > (unsigned long long)DENSE_MAX_FACTOR * (unsigned long long)nR)) return;
  }
  long long lo = 0x7FFFFFFFFFFFFFFFLL, hi = -0x7FFFFFFFFFFFFFFFLL - 1;
  for (int64_t i = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; i < nR; i += (int64_t)gridDim.x * BLOCK_THREADS) {
    const long long k = (long long)R[i];
    lo = k < lo ? k : lo; hi = k > hi ? k : hi;
  }
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, lo, d), b = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = a < lo ? a : lo; hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BLOCK_THREADS / 32; w++) { lo = smin[w] < lo ? smin[w] : lo; hi = smax[w] > hi ? smax[w] : hi; }
    atomicMin(&hdr->kmin, lo); atomicMax(&hdr->kmax, hi);
  }
}

__global__ void k_decide(TableHeader* hdr, int allow_dense) {
  const unsigned long long n = hdr->n_rows;
  if (n == 0 || !allow_dense || hdr->has_dups) return;             // duplicates proven by the sample: the direct-address build would only be aborted
  const unsigned long long range = (unsigned long long)hdr->kmax - (unsigned long long)hdr->kmin + 1ULL;   // wraps to 0 on the full 64-bit span
  if (range != 0 && range <= DENSE_MAX_FACTOR * n && range * 4 <= hdr->body_bytes && range <= 0xFFFFFFFFULL) {
    hdr->mode = MODE_DENSE; hdr->dense_range = range;
  }
}

// fallback_pass == 0: clear what the decided layout needs. fallback_pass == 1: only after a dense build hit a duplicate.
__global__ void __launch_bounds__(BLOCK_THREADS) k_clear(TableHeader* hdr, int4* body, int fallback_pass) {
  if (fallback_pass == 1) {
    if (!hdr->need_fallback) return;
  } else if (fallback_pass == 2) {
    if (hdr->mode != MODE_GROUP) return;
  }
  if (!fallback_pass && hdr->mode == MODE_HASH && hdr->has_dups) return;          // the sample already found duplicates: no inline attempt
  const bool dense = !fallback_pass && hdr->mode == MODE_DENSE;
  const unsigned long long bytes = dense ? ((hdr->dense_range * 4 + 15) & ~15ULL) : hdr->n_pairs * 64;
  const unsigned long long n16 = bytes / 16;
  const int4 ff = make_int4(-1, -1, -1, -1);
  for (unsigned long long i = blockIdx.x * (unsigned long long)BLOCK_THREADS + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * BLOCK_THREADS) body[i] = ff;
}

// Duplicate pre-check. Finding out that the build keys are not unique by building the inline table costs a whole aborted build
// (config 4: reorder 0.4 ms + clear + k_build_hash 0.67 ms before the grouped rebuild starts). Sixteen CTAs look at 4 096
// pseudo-random rows each first: equal keys at two different rows are proof of duplicates (no false positives), has_dups is set
// and the inline attempt is skipped; a miss only means the old path runs. With every key present 4 times (config 4) the samples
// hold ~12 such pairs. (One CTA with all 16 384 samples needs 192 KB of shared memory, and a launch with that carve-out costs
// 27 us even when the kernel exits at once; 48 KB per CTA costs 3 us.)
constexpr int DUPS_THREADS = 1024, DUPS_CTAS = 16, DUPS_SAMPLES = 4096, DUPS_SLOTS = 8192;
constexpr int64_t DUPS_MIN_ROWS = (int64_t)1 << 18;
__device__ __forceinline__ uint64_t sample_pos(uint32_t i, uint64_t n) { return __umul64hi(mix64((uint64_t)i + 0x9E3779B97F4A7C15ULL), n); }
template <typename K>
__global__ void __launch_bounds__(DUPS_THREADS) k_sample_dups(const K* __restrict__ R, int64_t nR, TableHeader* hdr) {
  __shared__ long long keys_sm[DUPS_SAMPLES];
  __shared__ unsigned short slots[DUPS_SLOTS];                   // sample id + 1, 0 = empty
  const uint32_t first = blockIdx.x * DUPS_SAMPLES;               // this CTA's samples: global ids [first, first + DUPS_SAMPLES)
  for (int s = threadIdx.x; s < DUPS_SLOTS; s += DUPS_THREADS) slots[s] = 0;
  for (int i = threadIdx.x; i < DUPS_SAMPLES; i += DUPS_THREADS) keys_sm[i] = (long long)R[sample_pos(first + i, (uint64_t)nR)];
  __syncthreads();
  bool dup = false;
  for (int i = threadIdx.x; i < DUPS_SAMPLES; i += DUPS_THREADS) {
    const long long key = keys_sm[i];
    uint32_t h = (uint32_t)(mix64((uint64_t)key) >> 40) & (DUPS_SLOTS - 1);
    for (;;) {
      const unsigned short old = atomicCAS(&slots[h], (unsigned short)0, (unsigned short)(i + 1));
      if (old == 0) break;
      const int j = old - 1;
      if (keys_sm[j] == key) { dup |= sample_pos(first + j, (uint64_t)nR) != sample_pos(first + i, (uint64_t)nR); break; }   // the same row drawn twice proves nothing
      h = (h + 1) & (DUPS_SLOTS - 1);
    }
  }
  // has_dups doubles as the NUMBER of sampled rows that met an equal key among their CTA's 4 096 samples (0 = none seen): with
  // rows spread at random it estimates the multiplicity of the build keys, m ~ 1 + pairs * n / (DUPS_CTAS * C(4096, 2))
  const int pairs = __syncthreads_count(dup);
  if (pairs && threadIdx.x == 0) atomicAdd(&hdr->has_dups, (uint32_t)pairs);
  long long lo = 0x7FFFFFFFFFFFFFFFLL, hi = -0x7FFFFFFFFFFFFFFFLL - 1;     // sampled key range: lets k_minmax skip its scan when it is already too wide
  for (int i = threadIdx.x; i < DUPS_SAMPLES; i += DUPS_THREADS) { const long long k = keys_sm[i]; lo = k < lo ? k : lo; hi = k > hi ? k : hi; }
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, lo, d), b = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = a < lo ? a : lo; hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&hdr->kmin, lo); atomicMax(&hdr->kmax, hi); }
}

// count_by_range (hjSetAllowDense(2), the default): a unique build whose keys are exactly [kmin, kmax] lets the count pass skip the
// table (a probe key matches iff it is in range) and moves the single lookup per row into the write pass. With the generic kernels
// it traded 0.9 ms of count for 1.0 ms of write on config 2 (2.10 vs 1.98 ms); with k_count_range / k_write_range (below) the step
// takes 1.84 ms instead of 1.99 ms.
__global__ void k_fallback_prepare(TableHeader* hdr, int count_by_range) {
  if (hdr->mode == MODE_DENSE && hdr->dense_filled != hdr->n_rows) hdr->need_fallback = 1;     // two rows stored into one slot: a duplicate key
  if (hdr->need_fallback) { hdr->mode = MODE_HASH; hdr->dense_range = 0; hdr->has_dups = 0; }
  else if (count_by_range && hdr->mode == MODE_DENSE && hdr->dense_range == hdr->n_rows) hdr->all_present = 1;
}

template <typename K, bool VEC>
__device__ __forceinline__ void load_vec_keys(const K* __restrict__ src, int64_t n, int64_t i0, uint64_t pol, K* key) {
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  if (VEC && i0 + KPV <= n) {
    int4 x = ld_stream_v4(src + i0, pol);
    memcpy(key, &x, 16);
  } else {
    #pragma unroll
    for (int e = 0; e < KPV; e++) key[e] = (i0 + e < n) ? src[i0 + e] : K(0);
  }
}

// same, 32-bit positions relative to a chunk of `n` rows
template <typename K, bool VEC>
__device__ __forceinline__ void load_vec_keys32(const K* __restrict__ src, uint32_t n, uint32_t i0, uint64_t pol, K* key) {
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  if (VEC && i0 + KPV <= n) {
    int4 x = ld_stream_v4(src + i0, pol);
    memcpy(key, &x, 16);
  } else {
    #pragma unroll
    for (int e = 0; e < KPV; e++) key[e] = (i0 + e < n) ? src[i0 + e] : K(0);
  }
}

template <typename K, bool VEC>
__global__ void __launch_bounds__(BLOCK_THREADS) k_build_dense(const K* __restrict__ R, int64_t nR, const uint32_t* __restrict__ payload, uint32_t row_base,
                                                               uint32_t* __restrict__ tab, TableHeader* hdr) {
  if (hdr->mode != MODE_DENSE) return;
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  const long long kmin = hdr->kmin;
  const uint64_t pol = policy_evict_first();
  // Plain stores, no atomics: whether two rows shared a slot is found afterwards by counting the slots taken (k_dense_verify, one
  // coalesced pass over a table that is still in L2: 15 us for 64 MB) — config 2's build 0.251 -> 0.219 ms.
  for (int64_t i0 = (blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x) * KPV; i0 < nR; i0 += (int64_t)gridDim.x * BLOCK_THREADS * KPV) {
    K key[KPV];
    load_vec_keys<K, VEC>(R, nR, i0, pol, key);
    #pragma unroll
    for (int e = 0; e < KPV; e++) {
      if (i0 + e < nR) tab[(unsigned long long)((long long)key[e] - kmin)] = payload ? payload[i0 + e] : row_base + (uint32_t)(i0 + e);
    }
  }
}
__global__ void __launch_bounds__(BLOCK_THREADS) k_dense_verify(const uint32_t* __restrict__ tab, TableHeader* hdr) {
  if (hdr->mode != MODE_DENSE) return;
  __shared__ unsigned long long red[33];
  const unsigned long long n = hdr->dense_range, n4 = n / 4;
  const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(tab);
  unsigned long long cnt = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)BLOCK_THREADS + threadIdx.x; i < n4; i += (unsigned long long)gridDim.x * BLOCK_THREADS) {
    const uint4 x = t4[i];
    cnt += (x.x != ROW_NONE) + (x.y != ROW_NONE) + (x.z != ROW_NONE) + (x.w != ROW_NONE);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) cnt += tab[n4 * 4 + threadIdx.x] != ROW_NONE;
  cnt = block_reduce_sum(cnt, red);
  if (threadIdx.x == 0 && cnt) atomicAdd(&hdr->dense_filled, cnt);
}

// Insert into the bucketised table: read the bucket once, CAS the first slot seen EMPTY; a failed CAS returns the
// occupant, which is final, so every earlier slot of the probe sequence has been compared with `key` by the time the
// insert lands -> duplicate detection is exact (the later of two equal keys always sees the earlier one).
// Measured alternatives on 2^28 i64 rows (build phase 15.0 ms): claiming slot 0 BLIND with CAS.128, four claims in flight per
// thread and 8 keys per ticket: 17.8 ms (1.4 CAS per key instead of 1.03); loading the home buckets of a vector up front (48
// registers instead of 32): k_build_hash 8.4 -> 9.4 ms. tools/membench4 (profiles/r1_membench4_insert_cost.jsonl): an L2-resident
// CAS.128 runs at 96 G/s, a 32-byte load at 272 G/s, load-then-CAS at 47 G/s, any 16-byte random write beyond L2 at 22 G/s.
template <typename K>
__device__ __forceinline__ bool insert_one(char* body, uint64_t n_pairs, K key, uint32_t row, const volatile uint32_t* has_dups) {
  using T = KeyTraits<K>;
  uint64_t pair; uint32_t half;
  T::home(key, n_pairs, pair, half);
  bool dup = false;
  for (uint32_t t = 0;; t++) {
    if (dup || (t >= 4 && *has_dups)) return true;             // duplicates known: this table is abandoned (grouped rebuild follows)
    char* bp = body + probe_bucket(pair, half, t, n_pairs) * 32;
    Bucket b = ld_bucket(bp);
    #pragma unroll
    for (int e = 0; e < T::SLOTS; e++) {
      if (slot_empty(b, e, key)) { if (slot_claim(bp, b, e, key, row)) return dup; }
      dup |= slot_match(b, e, key);
    }
  }
}

template <typename K, bool VEC>
__global__ void __launch_bounds__(BLOCK_THREADS) k_build_hash(const K* __restrict__ R, int64_t nR,
                                                              const uint32_t* __restrict__ payload, uint32_t row_base, char* body, TableHeader* hdr) {
  if (hdr->mode != MODE_HASH || hdr->has_dups) return;             // has_dups before the first insert: k_sample_dups found duplicates
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  const uint64_t n_pairs = hdr->n_pairs;
  const uint64_t pol = policy_evict_first();
  bool dup = false;
  __shared__ TicketQueue tq;
  const long long nblocks = (nR + (long long)BLOCK_THREADS * KPV - 1) / ((long long)BLOCK_THREADS * KPV);
  long long blk = ticket_first(&hdr->work[0], &tq);
  for (uint32_t it = 0; blk < nblocks; it++) {
    const long long pending = ticket_prefetch(&hdr->work[0]);
    const int64_t i0 = (blk * BLOCK_THREADS + threadIdx.x) * KPV;
    K key[KPV];
    load_vec_keys<K, VEC>(R, nR, i0, pol, key);
    #pragma unroll
    for (int e = 0; e < KPV; e++) {
      if (i0 + e < nR) {
        const uint32_t row = payload ? payload[i0 + e] : row_base + (uint32_t)(i0 + e);
        dup |= insert_one<K>(body, n_pairs, key[e], row, &hdr->has_dups);
      }
    }
    blk = ticket_advance(&tq, it, pending);
  }
  if (__ballot_sync(0xffffffffu, dup) && (threadIdx.x & 31) == 0) atomicOr(&hdr->has_dups, 1u);
}


// ---- grouped rebuild (runs only when the inline build found duplicate keys) --------------------------------
__global__ void k_group_prepare(TableHeader* hdr) {
  if (!hdr->has_dups || hdr->mode != MODE_HASH) return;
  const unsigned long long rows_bytes = (hdr->n_rows * 4 + 63) & ~63ULL;
  if (hdr->body_bytes < rows_bytes + 256) return;                  // cannot happen with hjTableBytes-sized workspaces
  hdr->mode = MODE_GROUP;
  hdr->n_pairs = (hdr->body_bytes - rows_bytes) / 64;
  hdr->rows_offset = hdr->n_pairs * 64;
  hdr->group_cursor = 0; hdr->n_groups = 0;
}

// find-or-insert the (sign-extended) key; returns the address of the slot's payload word
__device__ __forceinline__ unsigned long long* group_slot(char* body, uint64_t n_pairs, long long key, bool insert) {
  uint64_t pair; uint32_t half;
  group_home(key, n_pairs, pair, half);
  for (uint32_t t = 0;; t++) {
    char* bp = body + probe_bucket(pair, half, t, n_pairs) * 32;
    Bucket b = ld_bucket(bp);
    #pragma unroll
    for (int e = 0; e < 2; e++) {
      if ((uint32_t)b.w[2 * e + 1] == ROW_NONE) {
        if (!insert) return nullptr;
        const ulonglong2 old = atom_cas128(reinterpret_cast<ulonglong2*>(bp) + e, make_ulonglong2(~0ULL, ~0ULL), make_ulonglong2((unsigned long long)key, 0ULL));
        b.w[2 * e] = old.x; b.w[2 * e + 1] = old.y;
        if ((uint32_t)old.y == ROW_NONE) return reinterpret_cast<unsigned long long*>(bp) + 2 * e + 1;     // claimed: payload = 0
      }
      if (b.w[2 * e] == (unsigned long long)key) return reinterpret_cast<unsigned long long*>(bp) + 2 * e + 1;
    }
  }
}

template <typename K, bool VEC>
__global__ void __launch_bounds__(BLOCK_THREADS) k_group_count(const K* __restrict__ R, int64_t nR, char* body, TableHeader* hdr) {
  if (hdr->mode != MODE_GROUP) return;
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  const uint64_t n_pairs = hdr->n_pairs;
  const uint64_t pol = policy_evict_first();
  __shared__ TicketQueue tq;
  const long long nblocks = (nR + (long long)BLOCK_THREADS * KPV - 1) / ((long long)BLOCK_THREADS * KPV);
  long long blk = ticket_first(&hdr->work[1], &tq);
  for (uint32_t it = 0; blk < nblocks; it++) {
    const long long pending = ticket_prefetch(&hdr->work[1]);
    const int64_t i0 = (blk * BLOCK_THREADS + threadIdx.x) * KPV;
    K key[KPV];
    load_vec_keys<K, VEC>(R, nR, i0, pol, key);
    #pragma unroll
    for (int e = 0; e < KPV; e++)
      if (i0 + e < nR) atomicAdd(group_slot(body, n_pairs, (long long)key[e], true), 1ULL << 32);   // count lives in the high half
    blk = ticket_advance(&tq, it, pending);
  }
}

// hand every occupied slot a contiguous range of the row-id array: four slots per thread, ONE block scan per 1024 slots (the key
// counts and the number of occupied slots travel packed in one 64-bit word: both stay below 2^32 <= 2^40) + one atomic per CTA
constexpr int GROUP_OFF_ITEMS = 4;
__global__ void __launch_bounds__(BLOCK_THREADS) k_group_offsets(char* body, TableHeader* hdr) {
  if (hdr->mode != MODE_GROUP) return;
  __shared__ unsigned long long sm[33];
  __shared__ unsigned long long base_sm;
  constexpr unsigned long long LOW40 = (1ULL << 40) - 1;
  const unsigned long long n_slots = hdr->n_pairs * 4;
  for (unsigned long long s0 = (unsigned long long)blockIdx.x * BLOCK_THREADS * GROUP_OFF_ITEMS; s0 < n_slots;
       s0 += (unsigned long long)gridDim.x * BLOCK_THREADS * GROUP_OFF_ITEMS) {
    const unsigned long long s = s0 + (unsigned long long)threadIdx.x * GROUP_OFF_ITEMS;
    unsigned long long w[GROUP_OFF_ITEMS], packed = 0;
    #pragma unroll
    for (int i = 0; i < GROUP_OFF_ITEMS; i++) {
      w[i] = s + i < n_slots ? *(reinterpret_cast<const unsigned long long*>(body + (s + i) * 16) + 1) : ~0ULL;
      if ((uint32_t)w[i] != ROW_NONE) packed += (w[i] >> 32) + (1ULL << 40);
    }
    unsigned long long total, ex = block_exclusive_scan(packed, sm, &total);
    if (threadIdx.x == 0) {
      base_sm = (total & LOW40) ? atomicAdd(&hdr->group_cursor, total & LOW40) : 0ULL;
      if (total >> 40) atomicAdd(&hdr->n_groups, total >> 40);
    }
    __syncthreads();
    unsigned long long run = base_sm + (ex & LOW40);
    #pragma unroll
    for (int i = 0; i < GROUP_OFF_ITEMS; i++) {
      if ((uint32_t)w[i] != ROW_NONE) {                                // low half: start offset, advanced to `end` by the fill pass
        *(reinterpret_cast<unsigned long long*>(body + (s + i) * 16) + 1) = (w[i] & 0xFFFFFFFF00000000ULL) | (uint32_t)run;
        run += w[i] >> 32;
      }
    }
    __syncthreads();
  }
}

template <typename K, bool VEC>
__global__ void __launch_bounds__(BLOCK_THREADS) k_group_fill(const K* __restrict__ R, int64_t nR,
                                                              const uint32_t* __restrict__ payload, uint32_t row_base, char* body, TableHeader* hdr) {
  if (hdr->mode != MODE_GROUP) return;
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  const uint64_t n_pairs = hdr->n_pairs;
  uint32_t* rows = reinterpret_cast<uint32_t*>(body + hdr->rows_offset);
  const uint64_t pol = policy_evict_first();
  __shared__ TicketQueue tq;
  const long long nblocks = (nR + (long long)BLOCK_THREADS * KPV - 1) / ((long long)BLOCK_THREADS * KPV);
  long long blk = ticket_first(&hdr->work[2], &tq);
  for (uint32_t it = 0; blk < nblocks; it++) {
    const long long pending = ticket_prefetch(&hdr->work[2]);
    const int64_t i0 = (blk * BLOCK_THREADS + threadIdx.x) * KPV;
    K key[KPV];
    load_vec_keys<K, VEC>(R, nR, i0, pol, key);
    #pragma unroll
    for (int e = 0; e < KPV; e++) {
      if (i0 + e < nR) {
        unsigned long long* pay = group_slot(body, n_pairs, (long long)key[e], false);
        const unsigned long long old = atomicAdd(pay, 1ULL);      // low half is the write cursor of this key's range
        rows[(uint32_t)old] = payload ? payload[i0 + e] : row_base + (uint32_t)(i0 + e);
      }
    }
    blk = ticket_advance(&tq, it, pending);
  }
}

// kernels that may have nothing to do (their layout was not chosen) run on a bounded grid and loop: an idle launch then
// costs ~3 us instead of ~12 us for 16 K CTAs that only read the header and exit
constexpr int PERSIST_GRID = 148 * 8;
// Slice-ordered kernels walk the relation grid-stride, so the CTAs in flight cover ONE contiguous window of it (and one slice of
// the table). That only holds while every CTA of the grid is resident: a grid larger than one wave would sweep the table once per
// wave. resident_grid() = occupancy x SM count for the given kernel, cached.
// SM count and per-kernel occupancy are looked up once per (kernel, device) and cached; the caches are plain arrays indexed by device
// ordinal, filled idempotently, so concurrent first calls from several host threads are harmless.
constexpr int MAX_DEVICES = 64;
static int current_device() { int dev = 0; cudaGetDevice(&dev); return dev < 0 || dev >= MAX_DEVICES ? 0 : dev; }
static int num_sms() {
  static int cache[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (!cache[dev]) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); cache[dev] = n > 0 ? n : 148; }
  return cache[dev];
}
template <typename Kern>
static unsigned resident_grid(Kern kern, int64_t needed_blocks) {
  static int cache[MAX_DEVICES] = {0};                     // one array per kernel instantiation (the template makes it so)
  const int dev = current_device();
  if (!cache[dev]) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    cache[dev] = per_sm;
  }
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(needed_blocks, (int64_t)cache[dev] * num_sms()));
}

// ---- policy: process defaults (hjSet*), frozen into the table header by every build (hjBuildEx can state them per table) ----
static int g_allow_dense = 2;   // 0 hash only, 1 direct-address layout with the match cache, 2 (default) + count by range for gap-free unique key ranges
static int g_locality = 1;      // radix join for tables beyond L2 reach
static int g_dup_sample = 1;    // look at a sample of the build keys for duplicates before trying the inline (unique-key) layout
static int g_sparse = 1;        // hit lists for selective joins: 0 never, 1 sampled on the device, 2 always (unique layouts)
static int g_sliced = 1;       // slice-ordered inline layout for tables of 48 MB .. 1 GB of buckets (else: radix layout)
static int g_tma_count = 0;     // measured on C2: 4.52 ms with TMA-staged streams vs 1.22 ms with LDG/STG (profiles/README.md) -> off
void set_dup_sample(int on) { g_dup_sample = on; }
void set_sparse(int policy) { g_sparse = policy; }
void set_tma_count(int on) { g_tma_count = on; }
void set_allow_dense(int on) { g_allow_dense = on; }
void set_locality(int on) { g_locality = on; }
void set_sliced(int on) { g_sliced = on; }
uint32_t default_policy() {
  return (uint32_t)(g_allow_dense & 3) | (g_locality ? POLICY_RADIX : 0u) | ((uint32_t)(g_sparse & 3) << POLICY_SPARSE_SHIFT) | (g_dup_sample ? POLICY_DUP_SAMPLE : 0u) |
         (g_tma_count ? POLICY_TMA_COUNT : 0u) | (g_sliced ? 0u : POLICY_NO_SLICES);
}
// Grid of the direct-address probe kernels: 0 = one chunk per CTA, k = at most k resident waves striding over the chunks.
// Measured on config 2 (20 steps): count 1.231 / 1.248 / 1.233 / 1.231 ms and write 0.536 / 0.597 / 0.591 / 0.555 ms for k = 0 / 1 / 2 / 4.
static int g_count_waves = 2, g_write_waves = 0;
void set_dense_waves(int k) { g_count_waves = g_write_waves = k; }
template <typename Kern>
static unsigned dense_grid(Kern kern, int64_t nchunks, int waves) {
  if (waves <= 0) return (unsigned)std::min<int64_t>(nchunks, 16384);      // one chunk per CTA up to config 2's size, striding beyond
  return (unsigned)std::min<int64_t>(nchunks, (int64_t)waves * resident_grid(kern, nchunks));
}
static unsigned range_grid(int64_t nchunks) { return (unsigned)std::min<int64_t>(nchunks, 16384); }

// Small device -> host readbacks (table header, scratch counters, result size) land in a pinned block that belongs to the calling
// THREAD: two host threads driving two streams never share a landing zone.
static void* pinned_block() {
  thread_local void* p = nullptr;
  if (!p && cudaMallocHost(&p, 1024) != cudaSuccess) p = nullptr;
  return p;
}
cudaError_t readback(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t stream) {
  char* pin = reinterpret_cast<char*>(pinned_block());
  if (!pin || bytes > 1024) return cudaErrorMemoryAllocation;
  cudaError_t e = cudaMemcpyAsync(pin, dev_src, bytes, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) memcpy(host_dst, pin, bytes);
  return e;
}
static cudaError_t read_header(const void* table, TableHeader* out, cudaStream_t stream) { return readback(out, table, sizeof(TableHeader), stream); }

cudaError_t read_table_mode(const void* table, uint32_t* mode, uint32_t* all_present, cudaStream_t stream) {
  TableHeader h;
  cudaError_t e = read_header(table, &h, stream);
  if (e == cudaSuccess) { *mode = h.mode | (h.slice_bits ? 0x200u : 0u); if (all_present) *all_present = h.all_present; }
  return e;
}

__global__ void k_set_slices(TableHeader* hdr, uint32_t bits) { hdr->slice_bits = bits; }
// probe row ids of a slice-ordered copy that carries original indices (the reference's countRows states no row ids): out[j] = id of row idx[j]
__global__ void __launch_bounds__(BLOCK_THREADS) k_translate_rows(const uint32_t* __restrict__ idx, int64_t n, const uint32_t* __restrict__ payload, uint32_t row_base, uint32_t* __restrict__ out) {
  for (int64_t j = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; j < n; j += (int64_t)gridDim.x * BLOCK_THREADS) out[j] = payload ? payload[idx[j]] : row_base + idx[j];
}

template <typename K>
static cudaError_t launch_build(const K* R, int64_t nR, const uint32_t* payload, uint32_t row_base, char* body, int64_t body_bytes, TableHeader* hdr, int64_t pairs,
                                bool big, uint32_t policy, cudaStream_t stream) {
  constexpr int KPV = KeyTraits<K>::KEYS_PER_VEC;
  const int64_t threads = (nR + KPV - 1) / KPV;
  const unsigned grid = (unsigned)std::min<int64_t>(PERSIST_GRID, (threads + BLOCK_THREADS - 1) / BLOCK_THREADS);
  const unsigned clear_grid = (unsigned)std::min<int64_t>(148 * 16, (pairs * 4 + BLOCK_THREADS - 1) / BLOCK_THREADS);
  if (nR >= DUPS_MIN_ROWS && (policy & POLICY_DUP_SAMPLE)) k_sample_dups<K><<<DUPS_CTAS, DUPS_THREADS, 0, stream>>>(R, nR, hdr);
  if (nR > 0) k_minmax<K><<<(unsigned)std::min<int64_t>(148 * 8, grid), BLOCK_THREADS, 0, stream>>>(R, nR, hdr);
  k_decide<<<1, 1, 0, stream>>>(hdr, (int)(policy & POLICY_DENSE_MASK));
  // A table beyond L2 reach that did not get the direct-address layout is not built at all: one host look at the header (the only
  // sync in the build, and only for big tables), then the relation is radix-partitioned and the partitioned copy is the table (K7).
  if (big && nR > 0 && (policy & POLICY_RADIX)) {
    TableHeader h;
    cudaError_t e = read_header(hdr, &h, stream);
    if (e != cudaSuccess) return e;
    if (h.mode != MODE_DENSE) {
      // Heavy duplication (sampled multiplicity >= 32: the reference's published shape 1 has 100 rows per key) is output-bound, and there
      // the radix join's lockstep walk over long equal-key runs beats the grouped table's run expansion (10M x 10M -> 1e9 pairs: 4.2 vs
      // 5.4 ms); a few rows per key (config 4: 4) go the other way (6.65 vs 5.63 ms).
      const double multiplicity = 1.0 + (double)h.has_dups * (double)nR / ((double)DUPS_CTAS * 0.5 * DUPS_SAMPLES * (DUPS_SAMPLES - 1));
      if ((policy & POLICY_NO_SLICES) || !table_sliceable(nR, (int)sizeof(K)) || (h.has_dups && multiplicity >= 32.0))
        return radix_build(R, nR, (int)sizeof(K), payload, row_base, hdr, body, body_bytes, stream);
      // small enough to slice: reorder (key, row id) by table slice — of the grouped table when the sample already proved duplicates, else of
      // the inline table — then the ordinary build sequence runs over the copy (a duplicate the inline build meets still sends it to the
      // grouped layout, in the inline hash's slice order: correct, only less local for i32 keys)
      char* extra = body + table_part_bytes(nR, (int)sizeof(K));
      K* sk = reinterpret_cast<K*>(extra);
      uint32_t* sr = reinterpret_cast<uint32_t*>(extra + round_up(nR * (int64_t)sizeof(K), 256));
      uint32_t* so = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sr) + round_up(nR * 4, 256));
      void* ws = reinterpret_cast<char*>(so) + round_up(257 * 4, 256);
      const int bits = slice_bits_for(nR, (int)sizeof(K));
      e = slice_partition(R, payload, row_base, nR, (int)sizeof(K), bits, h.has_dups != 0, sk, sr, so, ws, slice_partition_workspace_bytes(nR, 8), stream);
      if (e != cudaSuccess) return e;
      k_set_slices<<<1, 1, 0, stream>>>(hdr, (uint32_t)bits);
      R = sk; payload = sr; row_base = 0;
    }
  }
  const bool vec = (reinterpret_cast<uintptr_t>(R) & 15) == 0;
  const int64_t need = (threads + BLOCK_THREADS - 1) / BLOCK_THREADS;
  k_clear<<<clear_grid, BLOCK_THREADS, 0, stream>>>(hdr, reinterpret_cast<int4*>(body), 0);
  if (nR == 0) return cudaGetLastError();
  if (policy & POLICY_DENSE_MASK) {
    if (vec) k_build_dense<K, true><<<resident_grid(k_build_dense<K, true>, need), BLOCK_THREADS, 0, stream>>>(R, nR, payload, row_base, reinterpret_cast<uint32_t*>(body), hdr);
    else     k_build_dense<K, false><<<resident_grid(k_build_dense<K, false>, need), BLOCK_THREADS, 0, stream>>>(R, nR, payload, row_base, reinterpret_cast<uint32_t*>(body), hdr);
    k_dense_verify<<<clear_grid, BLOCK_THREADS, 0, stream>>>(reinterpret_cast<const uint32_t*>(body), hdr);
    k_fallback_prepare<<<1, 1, 0, stream>>>(hdr, (policy & POLICY_DENSE_MASK) == 2);
    k_clear<<<clear_grid, BLOCK_THREADS, 0, stream>>>(hdr, reinterpret_cast<int4*>(body), 1);
  }
  if (vec) k_build_hash<K, true><<<resident_grid(k_build_hash<K, true>, need), BLOCK_THREADS, 0, stream>>>(R, nR, payload, row_base, body, hdr);
  else     k_build_hash<K, false><<<resident_grid(k_build_hash<K, false>, need), BLOCK_THREADS, 0, stream>>>(R, nR, payload, row_base, body, hdr);
  // duplicates found: rebuild in the grouped layout (every kernel below exits at once otherwise)
  k_group_prepare<<<1, 1, 0, stream>>>(hdr);
  k_clear<<<clear_grid, BLOCK_THREADS, 0, stream>>>(hdr, reinterpret_cast<int4*>(body), 2);
  if (vec) k_group_count<K, true><<<resident_grid(k_group_count<K, true>, need), BLOCK_THREADS, 0, stream>>>(R, nR, body, hdr);
  else     k_group_count<K, false><<<resident_grid(k_group_count<K, false>, need), BLOCK_THREADS, 0, stream>>>(R, nR, body, hdr);
  k_group_offsets<<<clear_grid, BLOCK_THREADS, 0, stream>>>(body, hdr);
  if (vec) k_group_fill<K, true><<<resident_grid(k_group_fill<K, true>, need), BLOCK_THREADS, 0, stream>>>(R, nR, payload, row_base, body, hdr);
  else     k_group_fill<K, false><<<resident_grid(k_group_fill<K, false>, need), BLOCK_THREADS, 0, stream>>>(R, nR, payload, row_base, body, hdr);
  return cudaGetLastError();
}

// (Measured and removed: raising cudaLimitPersistingL2CacheSize to its maximum, 79 MB, so that the table's L2::evict_last lines get
// the set-aside: config 2 1.84 -> 2.22 ms, forced hash 4.34 -> 6.47 ms. The streams lose more L2 than the table gains.)
cudaError_t build_table(const void* R, int64_t nR, int key_bytes, const uint32_t* payload, uint32_t row_base,
                        void* table, int64_t table_bytes_, uint32_t policy, cudaStream_t stream) {
  if (table_bytes_ < table_bytes(nR, key_bytes)) return cudaErrorInvalidValue;      // sized by hjTableBytes, nothing less
  const int64_t pairs = preferred_pairs(nR, key_bytes);
  if (pairs > (int64_t)1 << 32) return cudaErrorInvalidValue;
  TableHeader* hdr = reinterpret_cast<TableHeader*>(table);
  char* body = reinterpret_cast<char*>(table) + HEADER_BYTES;
  const int64_t part_bytes = table_part_bytes(nR, key_bytes);
  const bool big = table_is_big(nR, key_bytes);
  k_init_header<<<1, 1, 0, stream>>>(hdr, (uint32_t)key_bytes, (unsigned long long)nR, (unsigned long long)part_bytes, (unsigned long long)pairs, policy);
  if (key_bytes == 4) return launch_build<int32_t>((const int32_t*)R, nR, payload, row_base, body, table_bytes_ - HEADER_BYTES, hdr, pairs, big, policy, stream);
  return launch_build<int64_t>((const int64_t*)R, nR, payload, row_base, body, table_bytes_ - HEADER_BYTES, hdr, pairs, big, policy, stream);
}

// =========================================================================================================
// K2  count   (per probe row: unique build -> matched build row into the match cache; duplicates -> match count)
// =========================================================================================================
// element index of key k (0 .. KPT-1) of this thread inside a tile: (vec, thread, elem) order, coalesced 16-byte vectors
template <int KPV>
__device__ __forceinline__ int64_t elem_index(int64_t tile_base, int k) {
  return tile_base + ((int64_t)(k / KPV) * BLOCK_THREADS + threadIdx.x) * KPV + (k % KPV);
}

// u32 per probe row, same layout as the keys; the cache is padded to whole chunks so vector accesses never run out.
template <int KPV>
__device__ __forceinline__ void store_vec_u32(uint32_t* __restrict__ dst, int64_t i0, uint64_t pol, const uint32_t* m) {
  if (KPV == 4) st_stream_v4(dst + i0, make_int4(m[0], m[1], m[2], m[3]), pol);
  else st_stream_v2(dst + i0, make_uint2(m[0], m[1]), pol);
}
template <int KPV>
__device__ __forceinline__ void load_vec_u32(const uint32_t* __restrict__ src, int64_t i0, uint64_t pol, uint32_t* m) {
  if (KPV == 4) { int4 x = ld_stream_v4(src + i0, pol); m[0] = x.x; m[1] = x.y; m[2] = x.z; m[3] = x.w; }
  else { uint2 x = *reinterpret_cast<const uint2*>(src + i0); m[0] = x.x; m[1] = x.y; }
}

// first match of `key` (unique build): row id or ROW_NONE
template <typename K>
__device__ __forceinline__ uint32_t finish_probe_unique(const char* __restrict__ body, uint64_t n_pairs, K key, Bucket b) {
  using T = KeyTraits<K>;
  #pragma unroll
  for (int e = 0; e < T::SLOTS; e++) if (slot_match(b, e, key)) return slot_row(b, e, key);
  if (slot_empty(b, T::SLOTS - 1, key)) return ROW_NONE;          // home bucket not full: the sequence ends here (the common case)
  uint64_t pair; uint32_t half;                                    // rare: walk on; the home position is recomputed, not kept live
  T::home(key, n_pairs, pair, half);
  for (uint32_t t = 1;; t++) {
    b = ld_bucket(body + probe_bucket(pair, half, t, n_pairs) * 32);
    #pragma unroll
    for (int e = 0; e < T::SLOTS; e++) if (slot_match(b, e, key)) return slot_row(b, e, key);
    if (slot_empty(b, T::SLOTS - 1, key)) return ROW_NONE;
  }
}
template <typename K>
__device__ __forceinline__ const char* home_bucket(const char* __restrict__ body, uint64_t n_pairs, K key) {
  uint64_t pair; uint32_t half;
  KeyTraits<K>::home(key, n_pairs, pair, half);
  return body + probe_bucket(pair, half, 0, n_pairs) * 32;
}

// One instantiation per table layout (MODE); the host launches the one the table header names. Keeping them separate keeps registers
// per thread (and so occupancy) at what each layout needs. (The inline i32 variant holds four 32-byte buckets per thread: 64 registers,
// 4 CTAs per SM. Capped at 48 registers for a fifth CTA it spills 24 bytes and config 2 with sparse keys gets no faster: 2.69 -> 2.75 ms.)
template <typename K, bool VEC, uint32_t MODE>
__global__ void __launch_bounds__(BLOCK_THREADS) k_count(const K* __restrict__ S, int64_t nS, const char* __restrict__ body,
                                                         const TableHeader* __restrict__ hdr, uint32_t* __restrict__ mcache, uint32_t* __restrict__ run_start,
                                                         unsigned long long* __restrict__ chunk_totals, int64_t nchunks, unsigned long long* tickets,
                                                         const unsigned long long* __restrict__ sparse_flag, int semi) {
  using T = KeyTraits<K>;
  if (hdr->mode != MODE) return;
  if (MODE != MODE_GROUP && *sparse_flag) return;                  // selective join: k_count_sparse takes it
  if (MODE == MODE_DENSE && hdr->all_present) return;              // gap-free unique key range: k_count_range takes it
  constexpr int KPV = T::KEYS_PER_VEC, KPT = VECS_PER_THREAD * KPV, TILE = BLOCK_THREADS * KPT;
  __shared__ unsigned long long red[33];
  constexpr uint32_t mode = MODE;
  const uint64_t n_pairs = hdr->n_pairs;
  const long long kmin = hdr->kmin;
  const unsigned long long drange = hdr->dense_range;
  const uint64_t pol_s = policy_evict_first(), pol_t = policy_evict_last();
  constexpr int CHUNK_TILES = chunk_tiles((int)sizeof(K));
  __shared__ TicketQueue tq;
  // direct-address layout: chunks strided over the grid. Bucketised layouts: one resident wave taking chunks by ticket (fetched ahead).
  long long chunk = MODE == MODE_DENSE ? (long long)blockIdx.x : ticket_first(tickets, &tq);
  #pragma unroll 1
  for (uint32_t it = 0; chunk < nchunks; it++) {
  const long long pending = MODE == MODE_DENSE ? 0LL : ticket_prefetch(tickets);
  const int64_t chunk_base = chunk * (TILE * CHUNK_TILES);
  unsigned long long cnt = 0;

  #pragma unroll 1
  for (int tile = 0; tile < CHUNK_TILES; tile++) {
    const int64_t tile_base = chunk_base + (int64_t)tile * TILE;
    if (tile_base >= nS) break;
    if constexpr (mode != MODE_DENSE) {
      // bucketised layouts: one vector (KPV keys) at a time — its home buckets are in flight together, then each probe sequence
      // is finished. KPV x 32 bytes of bucket data in registers keeps the kernel at 5-6 CTAs per SM, which hides the serial
      // finish loops better than more loads per thread would (four keys at a time for i64: 39 -> 63 registers, count 6.2 -> 6.8 ms
      // on 2^28 x 2^28 rows).
      #pragma unroll 1
      for (int v = 0; v < VECS_PER_THREAD; v++) {
        const int64_t i0 = tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV;
        K kv[KPV]; uint32_t mv[KPV], sv[KPV]; Bucket b[KPV];
        load_vec_keys<K, VEC>(S, nS, i0, pol_s, kv);                        // rows past nS read as key 0: a harmless extra load
        #pragma unroll
        for (int e = 0; e < KPV; e++)
          b[e] = ld_bucket(mode == MODE_HASH ? home_bucket<K>(body, n_pairs, kv[e]) : home_bucket<int64_t>(body, n_pairs, (int64_t)kv[e]));
        #pragma unroll
        for (int e = 0; e < KPV; e++) {
          if constexpr (mode == MODE_HASH) { mv[e] = i0 + e < nS ? finish_probe_unique<K>(body, n_pairs, kv[e], b[e]) : ROW_NONE; cnt += (mv[e] != ROW_NONE); }
          else {
            // grouped: the slot's payload is (count << 32 | end of the key's row-id range). Count AND run start go to the scratch,
            // so the write pass neither reloads the keys nor probes the table again
            const unsigned long long pay = i0 + e < nS ? group_finish(body, n_pairs, (long long)kv[e], b[e]) : 0ULL;
            mv[e] = (uint32_t)(pay >> 32); sv[e] = (uint32_t)pay - mv[e];
            if (semi && mv[e]) mv[e] = 1;                                       // semi-join: a probe row counts once however many build rows match
            cnt += mv[e];
          }
        }
        store_vec_u32<KPV>(mcache, i0, pol_s, mv);
        if constexpr (mode == MODE_GROUP) store_vec_u32<KPV>(run_start, i0, pol_s, sv);
      }
    } else {
      // direct addressing: one 4-byte load per in-range key, all KPT in flight
      K key[KPT];
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_keys<K, VEC>(S, nS, tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV, pol_s, &key[v * KPV]);
      uint32_t m[KPT];
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        const unsigned long long off = (unsigned long long)((long long)key[k] - kmin);
        m[k] = (off < drange && elem_index<KPV>(tile_base, k) < nS) ? ld_keep_u32(reinterpret_cast<const uint32_t*>(body) + off, pol_t) : ROW_NONE;
      }
      #pragma unroll
      for (int k = 0; k < KPT; k++) cnt += (m[k] != ROW_NONE);
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) store_vec_u32<KPV>(mcache, tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV, pol_s, &m[v * KPV]);
    }
  }
  cnt = block_reduce_sum(cnt, red);
  if (threadIdx.x == 0) chunk_totals[chunk] = cnt;
  chunk = MODE == MODE_DENSE ? chunk + gridDim.x : ticket_advance(&tq, it, pending);
  }
}

// =========================================================================================================
// K2s  selective joins: hit lists instead of a match-cache word per probe row
// =========================================================================================================
// A match cache costs 4 bytes written and 4 bytes read per PROBE ROW whatever the selectivity (config 3: 8.6 GB of a 14 GB
// step for 10 % hits). When a sample of the probe keys says few rows hit, the count pass appends (matched build row, probe
// position) to a hit list of the chunk instead — the chunk's slice of the same scratch arrays, so capacity is never a
// question — and the write pass only copies lists to their final offsets. The decision is taken on the device
// (k_sample_hits -> counters[3]) and read uniformly by the kernels of both passes: no host round trip, and either path
// is correct for any selectivity (hjSetSparse(2) forces this one in the parity tests).
constexpr int SAMPLE_THREADS = 1024, SAMPLE_PER_THREAD = 2;   // 2048 samples: +-1 % on the hit fraction, one round of loads
constexpr int64_t SPARSE_MIN_ROWS = (int64_t)1 << 20;      // below this the sample costs more than it can save
constexpr unsigned SPARSE_MAX_PERCENT = 35;                // hit lists move 16 B per hit, the match cache 8 B per row

template <typename K>
__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample_hits(const K* __restrict__ S, int64_t nS, const char* __restrict__ body, const TableHeader* __restrict__ hdr,
                                                                unsigned long long* __restrict__ flag, int policy) {
  __shared__ unsigned long long red[33];
  const uint32_t mode = hdr->mode;
  if (mode == MODE_GROUP) return;                                // flag stays 0 (grouped: the cache holds counts)
  if (policy == 2) { if (threadIdx.x == 0) *flag = 1ULL; return; }
  const uint64_t n_pairs = hdr->n_pairs;
  const long long kmin = hdr->kmin;
  const unsigned long long drange = hdr->dense_range;
  unsigned long long hits = 0;
  #pragma unroll
  for (int s = 0; s < SAMPLE_PER_THREAD; s++) {
    const uint64_t j = __umul64hi(mix64((uint64_t)(threadIdx.x * SAMPLE_PER_THREAD + s) + 0x9E3779B97F4A7C15ULL), (uint64_t)nS);   // spread, not strided
    const K key = S[j];
    if (mode == MODE_DENSE) {
      const unsigned long long off = (unsigned long long)((long long)key - kmin);
      hits += off < drange && reinterpret_cast<const uint32_t*>(body)[off] != ROW_NONE;
    } else {
      hits += finish_probe_unique<K>(body, n_pairs, key, ld_bucket(home_bucket<K>(body, n_pairs, key))) != ROW_NONE;
    }
  }
  hits = block_reduce_sum(hits, red);
  if (threadIdx.x == 0) *flag = hits * 100 < (unsigned long long)SPARSE_MAX_PERCENT * SAMPLE_THREADS * SAMPLE_PER_THREAD ? 1ULL : 0ULL;
}

// Append the warp's hits among m[0..N) to the WARP's hit list (8-byte entries: matched build row, position of the probe row
// inside the chunk). Every warp owns a fixed sub-slice of the chunk's slice — as many entries as it has rows — so there is no
// cursor to share: no atomics, no block barrier. Inside the warp the order is thread-major (the order of a result is free,
// shared.cpp:168-171): each thread counts its hits, a shuffle scan ranks the threads, the hits are packed into the warp's
// shared-memory staging buffer with 32-bit addresses and leave as one coalesced run. The kernel is bound by instruction issue,
// not by memory (ncu on config 3: 2.9 IPC per SM at 365 warp instructions per tile for the first version, which ranked with one
// ballot per key slot and stored rows and positions to two arrays with 64-bit address arithmetic per slot), so the helpers
// below are PTX: nvcc turns the C++ forms into select chains, and wraps its own warp aggregation around a one-lane atomic.
__device__ __forceinline__ void count_hit(uint32_t& c, uint32_t row) {                  // c += row != ROW_NONE
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0xffffffff;\n @p add.u32 %0, %0, 1;\n}" : "+r"(c) : "r"(row));
}
__device__ __forceinline__ void stage_hit(uint32_t& saddr, uint32_t row, uint32_t pos) {   // if hit: *saddr++ = (row, pos)   (shared memory)
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0xffffffff;\n @p st.shared.v2.u32 [%0], {%1, %2};\n @p add.u32 %0, %0, 8;\n}"
               : "+r"(saddr) : "r"(row), "r"(pos) : "memory");
}
__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v) {               // shfl.up with its in-range predicate: 2 instructions per step
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1)
    asm volatile("{\n .reg .pred p;\n .reg .u32 t;\n shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n @p add.u32 %0, %0, t;\n}" : "+r"(v) : "r"(d));
  return v;
}
// returns the number of entries appended at wlist[0 ..)
template <int N>
__device__ __forceinline__ uint32_t append_hits(const uint32_t (&m)[N], const uint32_t (&pos)[N], uint2* __restrict__ wlist, uint2* stage) {
  const int lane = threadIdx.x & 31;
  uint32_t c = 0;
  #pragma unroll
  for (int k = 0; k < N; k++) count_hit(c, m[k]);
  const uint32_t inc = warp_inclusive_scan_u32(c);
  const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
  if (wtot == 0) return 0;                                         // warp-uniform
  uint32_t saddr = smem_addr(stage) + (inc - c) * 8;
  #pragma unroll
  for (int k = 0; k < N; k++) stage_hit(saddr, m[k], pos[k]);
  __syncwarp();
  for (uint32_t i = lane; i < wtot; i += 32) wlist[i] = stage[i];
  __syncwarp();                                                    // the staging buffer is rewritten by the next call
  return wtot;
}

constexpr int SPARSE_WARPS = BLOCK_THREADS / 32;
constexpr int PREFETCH_TILES = 2;

template <typename K, bool VEC, uint32_t MODE>
__global__ void __launch_bounds__(BLOCK_THREADS) k_count_sparse(const K* __restrict__ S, int64_t nS, const char* __restrict__ body, const TableHeader* __restrict__ hdr,
                                                                uint2* __restrict__ hit_list, uint32_t* __restrict__ warp_counts,
                                                                unsigned long long* __restrict__ chunk_totals, int64_t nchunks, unsigned long long* tickets,
                                                                const unsigned long long* __restrict__ sparse_flag) {
  using T = KeyTraits<K>;
  using UK = typename std::make_unsigned<K>::type;
  if (hdr->mode != MODE || !*sparse_flag) return;
  constexpr int KPV = T::KEYS_PER_VEC, KPT = VECS_PER_THREAD * KPV, TILE = BLOCK_THREADS * KPT;
  constexpr int CHUNK_TILES = chunk_tiles((int)sizeof(K)), CHUNK_ROWS = TILE * CHUNK_TILES, WARP_ROWS = CHUNK_ROWS / SPARSE_WARPS;
  __shared__ __align__(16) uint2 stage_sm[SPARSE_WARPS][32 * KPT];
  __shared__ uint32_t wc[2][SPARSE_WARPS];
  __shared__ long long ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2* stage = stage_sm[warp];
  const uint64_t n_pairs = hdr->n_pairs;
  // direct addressing in the key's own width: (UK)key - (UK)kmin < range is exact although it wraps (range <= 2^32 - 1, k_decide)
  const UK kmin = (UK)hdr->kmin, drange = (UK)hdr->dense_range;
  const uint32_t* __restrict__ tab = reinterpret_cast<const uint32_t*>(body);
  const uint64_t pol_s = policy_evict_first(), pol_t = policy_evict_last();
  int par = 0;
  // both layouts on one resident wave: direct-address strides over the chunks, the bucketised layout takes them by ticket (slice order)
  for (long long chunk = MODE == MODE_DENSE ? (long long)blockIdx.x : next_ticket(tickets, &ticket); chunk < nchunks;
       chunk = MODE == MODE_DENSE ? chunk + gridDim.x : next_ticket(tickets, &ticket), par ^= 1) {
    const int64_t chunk_base = chunk * CHUNK_ROWS;
    const uint32_t lim = (uint32_t)(nS - chunk_base < CHUNK_ROWS ? nS - chunk_base : CHUNK_ROWS);    // rows of this chunk
    const K* __restrict__ Sc = S + chunk_base;
    uint2* __restrict__ wlist = hit_list + chunk_base + warp * WARP_ROWS;
    uint32_t wcur = 0;
    #pragma unroll 1
    for (uint32_t tile_off = 0; tile_off < lim; tile_off += TILE) {
      if constexpr (MODE == MODE_DENSE) {
        K key[KPT];
        uint32_t m[KPT], pos[KPT];
        #pragma unroll
        for (int k = 0; k < KPT; k++) pos[k] = tile_off + ((k / KPV) * BLOCK_THREADS + threadIdx.x) * KPV + (k % KPV);
        if (VEC && tile_off + TILE <= lim) {                         // full tile (uniform): no per-row bounds tests
          {                                                          // pull the keys this thread reads two tiles from now into L2 (the kernel is latency-bound:
            const uint32_t ahead = tile_off + PREFETCH_TILES * TILE;   // ncu long-scoreboard stalls 10 warps per issue slot). A register double buffer measured
            const int64_t nb = ahead < CHUNK_ROWS ? chunk_base + ahead : (chunk + gridDim.x) * CHUNK_ROWS + (ahead - CHUNK_ROWS);   // slower: 48 registers, 5 CTAs per SM
            if (nb + TILE <= nS) {
              #pragma unroll
              for (int v = 0; v < VECS_PER_THREAD; v++) prefetch_l2(S + nb + (v * BLOCK_THREADS + threadIdx.x) * KPV);
            }
          }
          #pragma unroll
          for (int v = 0; v < VECS_PER_THREAD; v++) { const int4 x = ld_stream_v4(Sc + pos[v * KPV], pol_s); memcpy(&key[v * KPV], &x, 16); }
          #pragma unroll
          for (int k = 0; k < KPT; k++) { const UK off = (UK)key[k] - kmin; m[k] = off < drange ? ld_keep_u32(tab + off, pol_t) : ROW_NONE; }
        } else {
          #pragma unroll
          for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_keys32<K, VEC>(Sc, lim, pos[v * KPV], pol_s, &key[v * KPV]);
          #pragma unroll
          for (int k = 0; k < KPT; k++) { const UK off = (UK)key[k] - kmin; m[k] = (off < drange && pos[k] < lim) ? ld_keep_u32(tab + off, pol_t) : ROW_NONE; }
        }
        wcur += append_hits<KPT>(m, pos, wlist + wcur, stage);
      } else {
        #pragma unroll 1
        for (int v = 0; v < VECS_PER_THREAD; v++) {
          const uint32_t p0 = tile_off + (v * BLOCK_THREADS + threadIdx.x) * KPV;
          K kv[KPV]; uint32_t mv[KPV], pos[KPV]; Bucket b[KPV];
          load_vec_keys32<K, VEC>(Sc, lim, p0, pol_s, kv);
          #pragma unroll
          for (int e = 0; e < KPV; e++) b[e] = ld_bucket(home_bucket<K>(body, n_pairs, kv[e]));
          #pragma unroll
          for (int e = 0; e < KPV; e++) { pos[e] = p0 + e; mv[e] = p0 + e < lim ? finish_probe_unique<K>(body, n_pairs, kv[e], b[e]) : ROW_NONE; }
          wcur += append_hits<KPV>(mv, pos, wlist + wcur, stage);
        }
      }
    }
    if (lane == 0) { warp_counts[chunk * SPARSE_WARPS + warp] = wcur; wc[par][warp] = wcur; }
    __syncthreads();                                                 // wc is double-buffered: one barrier per chunk is enough
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      #pragma unroll
      for (int w = 0; w < SPARSE_WARPS; w++) t += wc[par][w];
      chunk_totals[chunk] = t;
    }
  }
}

// Write pass of the hit-list mode: one warp per chunk copies the chunk's eight warp lists to their final offsets (coalesced both
// ways) and turns probe positions into probe row ids (payload / row base, like k_write).
__global__ void __launch_bounds__(BLOCK_THREADS) k_write_sparse(const TableHeader* __restrict__ hdr, const uint2* __restrict__ hit_list, const uint32_t* __restrict__ warp_counts,
                                                                const unsigned long long* __restrict__ chunk_offsets, int64_t nchunks, int chunk_rows,
                                                                int32_t* __restrict__ outR, int32_t* __restrict__ outS,
                                                                const uint32_t* __restrict__ probe_payload, uint32_t probe_row_base,
                                                                const unsigned long long* __restrict__ sparse_flag) {
  if (!*sparse_flag || hdr->mode == MODE_GROUP) return;
  const int lane = threadIdx.x & 31;
  const uint64_t pol_s = policy_evict_first();
  const int64_t warps = (int64_t)gridDim.x * (BLOCK_THREADS / 32);
  const int warp_rows = chunk_rows / SPARSE_WARPS;
  for (int64_t chunk = (int64_t)blockIdx.x * (BLOCK_THREADS / 32) + (threadIdx.x >> 5); chunk < nchunks; chunk += warps) {
    const unsigned long long o = chunk_offsets[chunk];
    if (chunk_offsets[chunk + 1] == o) continue;                     // warp-uniform
    const uint32_t tot = (uint32_t)(chunk_offsets[chunk + 1] - o);
    const uint32_t cw = lane < SPARSE_WARPS ? warp_counts[chunk * SPARSE_WARPS + lane] : 0u;
    const uint32_t incl = warp_inclusive_scan_u32(cw);
    uint32_t end[SPARSE_WARPS];                                      // inclusive ends of the eight warp lists in the chunk's output range
    #pragma unroll
    for (int w = 0; w < SPARSE_WARPS; w++) end[w] = __shfl_sync(0xffffffffu, incl, w);
    const uint2* __restrict__ src = hit_list + chunk * chunk_rows;
    const uint32_t base = (uint32_t)(chunk * chunk_rows);
    // the eight lists are walked as ONE virtual list, four entries per lane in flight (one list at a time left the kernel latency-bound)
    constexpr int U = 4;
    for (uint32_t v0 = 0; v0 < tot; v0 += 32 * U) {
      uint2 e[U];
      #pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t v = v0 + u * 32 + lane;
        uint32_t start = 0, w = 0;
        #pragma unroll
        for (int x = 0; x < SPARSE_WARPS - 1; x++) if (v >= end[x]) { start = end[x]; w = x + 1; }
        e[u] = v < tot ? src[w * warp_rows + (v - start)] : make_uint2(0u, 0u);
      }
      #pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t v = v0 + u * 32 + lane;
        if (v < tot) {
          const uint32_t j = base + e[u].y;                                                // positions are chunk-relative
          if (outR) st_stream_u32(outR + o + v, e[u].x, pol_s);               // outR == nullptr: semi-join, probe rows only
          st_stream_u32(outS + o + v, probe_payload ? probe_payload[j] : probe_row_base + j, pol_s);
        }
      }
    }
  }
}

// Direct-address count with TMA-staged streams (i32 keys, 16-byte aligned probe column). Same result and same match-cache
// layout as k_count<K, true, MODE_DENSE>; the difference is HOW the two streams move: every warp owns a 256-row block of the
// tile (its own rows of both vectors), double-buffers it in shared memory with cp.async.bulk + its own mbarrier, and sends the
// match-cache words back with a bulk store. Only the table lookups go through LDG, so the L1TEX request path — the unit this
// kernel saturates — carries 2^28 lookups instead of 2^28 + 2 x 2^25 sector requests. No block-wide barrier in the loop.
__global__ void __launch_bounds__(BLOCK_THREADS) k_count_dense_tma(const int32_t* __restrict__ S, int64_t nS, const char* __restrict__ body,
                                                                   const TableHeader* __restrict__ hdr, uint32_t* __restrict__ mcache,
                                                                   unsigned long long* __restrict__ chunk_totals) {
  if (hdr->mode != MODE_DENSE || hdr->all_present) return;
  constexpr int KPV = 4, KPT = VECS_PER_THREAD * KPV, TILE = BLOCK_THREADS * KPT, WARPS = BLOCK_THREADS / 32;
  constexpr int CHUNK_TILES = chunk_tiles(4);
  constexpr uint32_t VEC_BYTES = 32 * KPV * 4;                              // one warp's rows of one vector: 512 bytes
  __shared__ __align__(128) int32_t keys_sm[2][WARPS][VECS_PER_THREAD][32 * KPV];
  __shared__ __align__(128) uint32_t m_sm[2][WARPS][VECS_PER_THREAD][32 * KPV];
  __shared__ __align__(8) unsigned long long bars[2][WARPS];
  __shared__ unsigned long long red[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long kmin = hdr->kmin;
  const unsigned long long drange = hdr->dense_range;
  const uint64_t pol_s = policy_evict_first(), pol_t = policy_evict_last();
  const int64_t chunk_base = (int64_t)blockIdx.x * (TILE * CHUNK_TILES);
  if (lane == 0) { mbar_init(&bars[0][warp], 1); mbar_init(&bars[1][warp], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  // rows of this warp in tile `t`: vector v covers [tile_base + (v * 256 + warp * 32) * 4, + 128 rows)
  auto issue = [&](int t, int buf) {
    const int64_t tile_base = chunk_base + (int64_t)t * TILE;
    if (lane == 0) {
      mbar_expect_tx(&bars[buf][warp], VECS_PER_THREAD * VEC_BYTES);
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++)
        tma_load_1d(keys_sm[buf][warp][v], S + tile_base + (int64_t)(v * BLOCK_THREADS + warp * 32) * KPV, VEC_BYTES, &bars[buf][warp], pol_s);
    }
  };
  // a tile is "full" when all its rows exist; the (single) ragged tile at the end of the relation takes the LDG path
  auto full = [&](int t) { return chunk_base + (int64_t)(t + 1) * TILE <= nS; };

  unsigned long long cnt = 0;
  uint32_t phase[2] = {0, 0};
  if (full(0)) issue(0, 0);
  #pragma unroll 1
  for (int t = 0; t < CHUNK_TILES; t++) {
    const int64_t tile_base = chunk_base + (int64_t)t * TILE;
    if (tile_base >= nS) break;
    const int buf = t & 1;
    if (t + 1 < CHUNK_TILES && full(t + 1)) issue(t + 1, buf ^ 1);        // prefetch the next tile's keys
    int32_t key[KPT];
    if (full(t)) {
      mbar_wait(&bars[buf][warp], phase[buf]); phase[buf] ^= 1;
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) { const int4 x = *reinterpret_cast<const int4*>(&keys_sm[buf][warp][v][lane * KPV]); memcpy(&key[v * KPV], &x, 16); }
    } else {
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_keys<int32_t, true>(S, nS, tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV, pol_s, &key[v * KPV]);
    }
    uint32_t m[KPT];
    #pragma unroll
    for (int k = 0; k < KPT; k++) {
      const unsigned long long off = (unsigned long long)((long long)key[k] - kmin);
      m[k] = (off < drange && elem_index<KPV>(tile_base, k) < nS) ? ld_keep_u32(reinterpret_cast<const uint32_t*>(body) + off, pol_t) : ROW_NONE;
    }
    #pragma unroll
    for (int k = 0; k < KPT; k++) cnt += (m[k] != ROW_NONE);
    // match-cache words of this warp leave through shared memory + one bulk store per vector (the cache is padded to whole chunks)
    tma_store_wait_read<VECS_PER_THREAD>();                               // the stores that last read m_sm[buf] (two tiles ago) are done reading
    __syncwarp();
    #pragma unroll
    for (int v = 0; v < VECS_PER_THREAD; v++) *reinterpret_cast<uint4*>(&m_sm[buf][warp][v][lane * KPV]) = make_uint4(m[v * KPV], m[v * KPV + 1], m[v * KPV + 2], m[v * KPV + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++)
        tma_store_1d(mcache + tile_base + (int64_t)(v * BLOCK_THREADS + warp * 32) * KPV, m_sm[buf][warp][v], VEC_BYTES, pol_s);
    }
  }
  tma_store_wait_read<0>();
  cnt = block_reduce_sum(cnt, red);
  if (threadIdx.x == 0) chunk_totals[blockIdx.x] = cnt;
}

// =========================================================================================================
// K2r / K4r  count by range (hjSetAllowDense(2)): a unique build whose keys are exactly [kmin, kmax] (direct-address layout with
// every slot taken; k_fallback_prepare sets all_present) lets the count pass skip the table — a probe key matches iff it is in
// range — and drop the match cache; the write pass does the one lookup per matching row. The ranking of a tile's hits needs only
// the range test, so in the write pass the ballots and the block barrier run while the lookups are in flight.
// =========================================================================================================
template <typename K, bool VEC>
__global__ void __launch_bounds__(BLOCK_THREADS) k_count_range(const K* __restrict__ S, int64_t nS, const TableHeader* __restrict__ hdr,
                                                               unsigned long long* __restrict__ chunk_totals, uint32_t* __restrict__ warp_first, int64_t nchunks,
                                                               const unsigned long long* __restrict__ sparse_flag) {
  using T = KeyTraits<K>;
  using UK = typename std::make_unsigned<K>::type;
  if (hdr->mode != MODE_DENSE || !hdr->all_present || *sparse_flag) return;     // few rows hit: hit lists read the keys once, this path twice
  constexpr int KPV = T::KEYS_PER_VEC, CHUNK_ROWS = chunk_keys((int)sizeof(K)), NV = CHUNK_ROWS / (BLOCK_THREADS * KPV);   // vectors per thread and chunk: 16 (i32) / 2 (i64)
  constexpr int UNROLL = NV < 8 ? NV : 8;
  constexpr int WARPS = BLOCK_THREADS / 32;
  __shared__ uint32_t wsum[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const UK kmin = (UK)hdr->kmin, drange = (UK)hdr->dense_range;
  const uint64_t pol_s = policy_evict_first();
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t chunk_base = chunk * CHUNK_ROWS;
    const uint32_t lim = (uint32_t)(nS - chunk_base < CHUNK_ROWS ? nS - chunk_base : CHUNK_ROWS);
    const K* __restrict__ Sc = S + chunk_base;
    uint32_t cnt = 0;                                          // hits among THIS thread's elements — the same elements it has in k_write_range
    if (VEC && lim == CHUNK_ROWS) {
      #pragma unroll 1
      for (int v0 = 0; v0 < NV; v0 += UNROLL) {
        int4 x[UNROLL];
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) x[u] = ld_stream_v4(Sc + ((v0 + u) * BLOCK_THREADS + threadIdx.x) * KPV, pol_s);
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) {
          K key[KPV]; memcpy(key, &x[u], 16);
          #pragma unroll
          for (int e = 0; e < KPV; e++) cnt += ((UK)key[e] - kmin) < drange;
        }
      }
    } else {
      for (int v = 0; v < NV; v++) {
        #pragma unroll
        for (int e = 0; e < KPV; e++) { const uint32_t i = (uint32_t)(v * BLOCK_THREADS + threadIdx.x) * KPV + e; cnt += i < lim && ((UK)Sc[i] - kmin) < drange; }
      }
    }
    // per chunk: its total (scanned by K3 into the chunk's first output element) and, per warp, the first output element of the warp's
    // hits relative to the chunk — with it every warp of k_write_range owns a contiguous output run and needs no block barrier
    cnt = warp_reduce_sum(cnt);
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    if (lane == 0) {
      uint32_t first = 0, total = 0;
      #pragma unroll
      for (int w = 0; w < WARPS; w++) { const uint32_t x = wsum[w]; first += w < warp ? x : 0u; total += x; }
      warp_first[chunk * WARPS + warp] = first;
      if (warp == 0) chunk_totals[chunk] = total;
    }
    __syncthreads();                                           // wsum is reused by the next chunk
  }
}

template <typename K, bool VEC>
__global__ void __launch_bounds__(BLOCK_THREADS) k_write_range(const K* __restrict__ S, int64_t nS, const char* __restrict__ body, const TableHeader* __restrict__ hdr,
                                                               const unsigned long long* __restrict__ chunk_offsets, const uint32_t* __restrict__ warp_first, int64_t nchunks,
                                                               int32_t* __restrict__ outR, int32_t* __restrict__ outS,
                                                               const uint32_t* __restrict__ probe_payload, uint32_t probe_row_base,
                                                               const unsigned long long* __restrict__ sparse_flag) {
  using T = KeyTraits<K>;
  using UK = typename std::make_unsigned<K>::type;
  if (hdr->mode != MODE_DENSE || !hdr->all_present || *sparse_flag) return;
  constexpr int KPV = T::KEYS_PER_VEC, KPT = VECS_PER_THREAD * KPV, TILE = BLOCK_THREADS * KPT, WARPS = BLOCK_THREADS / 32;
  constexpr int CHUNK_ROWS = chunk_keys((int)sizeof(K));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const UK kmin = (UK)hdr->kmin, drange = (UK)hdr->dense_range;
  const uint32_t* __restrict__ tab = reinterpret_cast<const uint32_t*>(body);
  const uint64_t pol_s = policy_evict_first(), pol_t = policy_evict_last();
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    unsigned long long o = chunk_offsets[chunk];
    if (chunk_offsets[chunk + 1] == o) continue;                               // nothing to emit for this chunk (uniform)
    o += warp_first[chunk * WARPS + warp];                                     // this warp's own output run (k_count_range): no block barrier below
    const int64_t chunk_base = chunk * CHUNK_ROWS;
    const uint32_t lim = (uint32_t)(nS - chunk_base < CHUNK_ROWS ? nS - chunk_base : CHUNK_ROWS);
    const K* __restrict__ Sc = S + chunk_base;
    #pragma unroll 1
    for (uint32_t tile_off = 0; tile_off < lim; tile_off += TILE) {
      K key[KPT]; uint32_t pos[KPT], m[KPT]; bool hit[KPT];
      #pragma unroll
      for (int k = 0; k < KPT; k++) pos[k] = tile_off + ((k / KPV) * BLOCK_THREADS + threadIdx.x) * KPV + (k % KPV);
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_keys32<K, VEC>(Sc, lim, pos[v * KPV], pol_s, &key[v * KPV]);
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        const UK off = (UK)key[k] - kmin;
        hit[k] = off < drange && pos[k] < lim;
        m[k] = hit[k] ? ld_keep_u32(tab + off, pol_t) : ROW_NONE;              // in flight while the hits are ranked below
      }
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        const unsigned mask = __ballot_sync(0xffffffffu, hit[k]);
        if (hit[k]) {
          const unsigned long long dst = o + __popc(mask & lt);
          const uint32_t j = (uint32_t)chunk_base + pos[k];
          if (outR) st_stream_u32(outR + dst, m[k], pol_s);
          st_stream_u32(outS + dst, probe_payload ? probe_payload[j] : probe_row_base + j, pol_s);
        }
        o += __popc(mask);
      }
    }
  }
}

// =========================================================================================================
// K3  scan of chunk totals -> exclusive chunk offsets, total at [nchunks]      (replaces join_v1.mlir:371-420)
// =========================================================================================================
// Two short launches: every CTA scans 16 384 totals in place (thread-contiguous 128-byte runs, one block scan), publishing its
// block total; the second launch (only when there is more than one block) adds the preceding blocks' totals. A single CTA
// looping over the array took 61 us for config 3's 65 536 chunks and would take ~1 ms for 2^30 i64 rows (1 M chunks).
constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 16, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(unsigned long long* __restrict__ t, int64_t n, unsigned long long* __restrict__ block_sums, unsigned long long* __restrict__ total_out) {
  __shared__ unsigned long long sm[33];
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long v[SCAN_ITEMS], sum = 0;
  if (i0 + SCAN_ITEMS <= n) {
    #pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e += 2) { const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(t + i0 + e); v[e] = x.x; v[e + 1] = x.y; }
  } else {
    #pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e++) v[e] = (i0 + e < n) ? t[i0 + e] : 0ULL;
  }
  #pragma unroll
  for (int e = 0; e < SCAN_ITEMS; e++) sum += v[e];
  unsigned long long total, acc = block_exclusive_scan(sum, sm, &total);
  #pragma unroll
  for (int e = 0; e < SCAN_ITEMS; e++) { const unsigned long long x = v[e]; v[e] = acc; acc += x; }
  if (i0 + SCAN_ITEMS <= n) {
    #pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e += 2) *reinterpret_cast<ulonglong2*>(t + i0 + e) = make_ulonglong2(v[e], v[e + 1]);
  } else {
    #pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e++) if (i0 + e < n) t[i0 + e] = v[e];
  }
  if (threadIdx.x == 0) { block_sums[blockIdx.x] = total; if (gridDim.x == 1) { t[n] = total; if (total_out) *total_out = total; } }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(unsigned long long* __restrict__ t, int64_t n, const unsigned long long* __restrict__ block_sums, unsigned long long* __restrict__ total_out) {
  __shared__ unsigned long long sm[33];
  unsigned long long mine = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += SCAN_THREADS) mine += block_sums[b];
  const unsigned long long base = block_reduce_sum(mine, sm);
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  if (blockIdx.x > 0) {
    #pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e++) if (i0 + e < n) t[i0 + e] += base;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { t[n] = base + block_sums[blockIdx.x]; if (total_out) *total_out = t[n]; }
}
// t[0 .. n) := exclusive prefix, t[n] := total (also stored at *total_out when given). block_sums: ceil(n / 16 384) entries of scratch.
void launch_scan(unsigned long long* t, int64_t n, unsigned long long* block_sums, unsigned long long* total_out, cudaStream_t stream) {
  const unsigned nb = (unsigned)std::max<int64_t>(1, (n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  k_scan_blocks<<<nb, SCAN_THREADS, 0, stream>>>(t, n, block_sums, total_out);
  if (nb > 1) k_scan_add<<<nb, SCAN_THREADS, 0, stream>>>(t, n, block_sums, total_out);
}

// The count pass looks at the table header first (one 256-byte readback: a stream sync) and launches exactly the kernels of the layout
// it finds, under the policy the table was BUILT with. Only the hit-list decision stays on the device (k_sample_hits -> CTR_SPARSE): for
// the unique layouts the match-cache / range kernel and the hit-list kernel are both queued and one of them exits at once.
cudaError_t count_rows_async(const void* S, int64_t nS, int key_bytes, const void* table, void* scratch,
                             bool carry_rows, const uint32_t* probe_payload, uint32_t probe_row_base, bool semi, cudaStream_t stream) {
  ScratchView sv = scratch_view(scratch, nS, key_bytes);
  const TableHeader* hdr = reinterpret_cast<const TableHeader*>(table);
  const char* body = reinterpret_cast<const char*>(table) + HEADER_BYTES;
  TableHeader h;
  { cudaError_t e = read_header(table, &h, stream); if (e != cudaSuccess) return e; }
  if (h.magic != HJ_MAGIC || (int)h.key_bytes != key_bytes) return cudaErrorInvalidValue;          // never built, or built for the other key width
  { cudaError_t e = cudaMemsetAsync(sv.counters, 0, SCRATCH_COUNTERS * sizeof(unsigned long long), stream); if (e != cudaSuccess) return e; }
  if (carry_rows) { cudaError_t e = cudaMemsetAsync(sv.counters + CTR_CARRIED, 1, 1, stream); if (e != cudaSuccess) return e; }
  if (semi) { cudaError_t e = cudaMemsetAsync(sv.counters + CTR_SEMI, 1, 1, stream); if (e != cudaSuccess) return e; }
  unsigned long long* total_out = sv.counters + CTR_TOTAL;
  if (h.mode == MODE_RADIX)
    return radix_count(S, nS, key_bytes, h, body, sv.radix, sv.chunk_offsets, sv.scan_sums, sv.counters + CTR_TICKET_RADIX, total_out, sv.mcache, sv.counters + CTR_RADIX_MULTI,
                       carry_rows, probe_payload, probe_row_base, semi, stream);
  if (h.slice_bits && (h.mode == MODE_HASH || h.mode == MODE_GROUP) && nS > 0) {       // slice-ordered table: the probe passes walk a slice-ordered copy
    const SliceArea sa = slice_area(sv.radix, nS, key_bytes);
    cudaError_t e = slice_partition(S, carry_rows ? probe_payload : nullptr, carry_rows ? probe_row_base : 0u, nS, key_bytes, (int)h.slice_bits, h.mode == MODE_GROUP,
                                    sa.keys, sa.rows, sa.offsets, sa.ws, sa.ws_bytes, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(sv.counters + CTR_SLICED, 1, 1, stream);
    if (e != cudaSuccess) return e;
    S = sa.keys;
  }
  const unsigned long long* sparse_flag = sv.counters + CTR_SPARSE;
  const bool tma = (h.policy & POLICY_TMA_COUNT) != 0;
  const int sparse_policy = (tma || h.mode == MODE_GROUP) ? 0 : (int)((h.policy >> POLICY_SPARSE_SHIFT) & 3);
  const bool by_range = h.mode == MODE_DENSE && h.all_present;
  if (nS > 0 && (sparse_policy == 2 || (sparse_policy == 1 && nS >= SPARSE_MIN_ROWS))) {
    if (key_bytes == 4) k_sample_hits<int32_t><<<1, SAMPLE_THREADS, 0, stream>>>((const int32_t*)S, nS, body, hdr, sv.counters + CTR_SPARSE, sparse_policy);
    else                k_sample_hits<int64_t><<<1, SAMPLE_THREADS, 0, stream>>>((const int64_t*)S, nS, body, hdr, sv.counters + CTR_SPARSE, sparse_policy);
  }
  if (sv.nchunks > 0) {
    const bool vec = (reinterpret_cast<uintptr_t>(S) & 15) == 0;
#define HJ_LAUNCH_COUNT(K, V) \
    if (h.mode == MODE_GROUP) k_count<K, V, MODE_GROUP><<<resident_grid(k_count<K, V, MODE_GROUP>, sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.mcache, sv.run_start, sv.chunk_offsets, sv.nchunks, sv.counters + CTR_TICKET_GROUP, sparse_flag, semi ? 1 : 0); \
    else if (h.mode == MODE_HASH) { \
      k_count<K, V, MODE_HASH><<<resident_grid(k_count<K, V, MODE_HASH>, sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.mcache, sv.run_start, sv.chunk_offsets, sv.nchunks, sv.counters + CTR_TICKET_HASH, sparse_flag, 0);  \
      if (sparse_policy) k_count_sparse<K, V, MODE_HASH><<<resident_grid(k_count_sparse<K, V, MODE_HASH>, sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.hit_list, sv.warp_counts, sv.chunk_offsets, sv.nchunks, sv.counters + CTR_TICKET_SPARSE, sparse_flag); \
    } else { \
      if (by_range) k_count_range<K, V><<<range_grid(sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, hdr, sv.chunk_offsets, sv.mcache, sv.nchunks, sparse_flag); \
      else if (tma && sizeof(K) == 4 && V) k_count_dense_tma<<<(unsigned)sv.nchunks, BLOCK_THREADS, 0, stream>>>((const int32_t*)S, nS, body, hdr, sv.mcache, sv.chunk_offsets); \
      else k_count<K, V, MODE_DENSE><<<dense_grid(k_count<K, V, MODE_DENSE>, sv.nchunks, g_count_waves), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.mcache, sv.run_start, sv.chunk_offsets, sv.nchunks, sv.counters, sparse_flag, 0); \
      if (sparse_policy) k_count_sparse<K, V, MODE_DENSE><<<resident_grid(k_count_sparse<K, V, MODE_DENSE>, sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.hit_list, sv.warp_counts, sv.chunk_offsets, sv.nchunks, sv.counters + CTR_TICKET_SPARSE, sparse_flag); \
    }
    if (key_bytes == 4) { if (vec) { HJ_LAUNCH_COUNT(int32_t, true) } else { HJ_LAUNCH_COUNT(int32_t, false) } }
    else                { if (vec) { HJ_LAUNCH_COUNT(int64_t, true) } else { HJ_LAUNCH_COUNT(int64_t, false) } }
#undef HJ_LAUNCH_COUNT
  }
  launch_scan(sv.chunk_offsets, sv.nchunks, sv.scan_sums, total_out, stream);
  return cudaGetLastError();
}

// =========================================================================================================
// K4  write
// =========================================================================================================
// Two instantiations (GROUPED = false: direct-address / inline layouts, true: grouped layout); the write launch queues both
// on a bounded grid and the one that does not match the header exits at once (registers: 40 vs 64 per thread).
template <typename K, bool VEC, bool GROUPED>
__global__ void __launch_bounds__(BLOCK_THREADS) k_write(const K* __restrict__ S, int64_t nS, const char* __restrict__ body,
                                                         const TableHeader* __restrict__ hdr, const uint32_t* __restrict__ mcache, const uint32_t* __restrict__ run_start,
                                                         const unsigned long long* __restrict__ chunk_offsets, int64_t nchunks, unsigned long long* tickets,
                                                         int32_t* __restrict__ outR, int32_t* __restrict__ outS,
                                                         const uint32_t* __restrict__ probe_payload, uint32_t probe_row_base,
                                                         const unsigned long long* __restrict__ sparse_flag) {
  if ((hdr->mode == MODE_GROUP) != GROUPED) return;
  if (!GROUPED && *sparse_flag) return;                              // hit-list mode: k_write_sparse takes it
  if (!GROUPED && hdr->all_present) return;                          // count-by-range mode: k_write_range takes it
  using T = KeyTraits<K>;
  constexpr int KPV = T::KEYS_PER_VEC, KPT = VECS_PER_THREAD * KPV, TILE = BLOCK_THREADS * KPT;
  auto probe_row = [&](int64_t j) -> uint32_t { return probe_payload ? probe_payload[j] : probe_row_base + (uint32_t)j; };
  __shared__ uint32_t warp_totals[2][BLOCK_THREADS / 32];
  __shared__ unsigned long long scan_sm[33];
  constexpr bool dups = GROUPED;
  const uint64_t pol_s = policy_evict_first();
  constexpr int CHUNK_TILES = chunk_tiles((int)sizeof(K));
  __shared__ long long ticket;
  #pragma unroll 1
  for (long long chunk = GROUPED ? next_ticket(tickets, &ticket) : (long long)blockIdx.x; chunk < nchunks;
       chunk = GROUPED ? next_ticket(tickets, &ticket) : chunk + gridDim.x) {
  const int64_t chunk_base = chunk * (TILE * CHUNK_TILES);
  unsigned long long out_base = chunk_offsets[chunk];
  if (chunk_offsets[chunk + 1] == out_base) continue;                        // nothing to emit for this chunk (uniform)

  #pragma unroll 1
  for (int tile = 0; tile < CHUNK_TILES; tile++) {
    const int64_t tile_base = chunk_base + (int64_t)tile * TILE;
    if (tile_base >= nS) break;
    uint32_t m[KPT];
    #pragma unroll
    for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_u32<KPV>(mcache, tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV, pol_s, &m[v * KPV]);

    if constexpr (!dups) {
      // unique build: the cache already holds the build row. Output order is free (the result is a multiset,
      // shared.cpp:168-171), so every warp owns one contiguous output range and compacts with ballots: no staging
      // buffer, one barrier per tile, and each store instruction writes one contiguous run of up to 128 bytes.
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      unsigned mask[KPT];
      uint32_t wtotal = 0;
      #pragma unroll
      for (int k = 0; k < KPT; k++) { mask[k] = __ballot_sync(0xffffffffu, m[k] != ROW_NONE); wtotal += __popc(mask[k]); }
      uint32_t* wt = warp_totals[tile & 1];
      if (lane == 0) wt[warp] = wtotal;
      __syncthreads();
      uint32_t wbase = 0, ttotal = 0;
      #pragma unroll
      for (int w = 0; w < BLOCK_THREADS / 32; w++) { const uint32_t x = wt[w]; wbase += w < warp ? x : 0u; ttotal += x; }
      unsigned long long o = out_base + wbase;
      const unsigned lt = (1u << lane) - 1u;
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        if (m[k] != ROW_NONE) {
          const int64_t j = elem_index<KPV>(tile_base, k);
          const unsigned long long dst = o + __popc(mask[k] & lt);
          if (outR) st_stream_u32(outR + dst, m[k], pol_s);
          st_stream_u32(outS + dst, probe_row(j), pol_s);
        }
        o += __popc(mask[k]);
      }
      out_base += ttotal;
    } else {
      // grouped layout: the cache holds per-row match counts. Every warp owns one contiguous output range, laid out key slot
      // by key slot (all lanes' runs of slot 0, then slot 1, ...). The warp expands its runs cooperatively: output element p of
      // a slot is found by a shuffle binary search over the lanes' inclusive counts, so both result columns leave as full
      // 128-byte lines and the row ids of a run are read with neighbouring lanes on neighbouring addresses.
      const uint32_t* __restrict__ rows = reinterpret_cast<const uint32_t*>(body + hdr->rows_offset);
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      uint32_t incl[KPT], start[KPT], prow[KPT];
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_u32<KPV>(run_start, tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV, pol_s, &start[v * KPV]);
      unsigned long long wtotal = 0;
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        incl[k] = warp_inclusive_scan(m[k]);                                  // a probe row has < 2^32 matches; a warp slot total is kept in 64 bits below
        prow[k] = m[k] ? probe_row(elem_index<KPV>(tile_base, k)) : 0u;       // the run start came from the count pass (no key reload, no second probe)
        wtotal += __shfl_sync(0xffffffffu, incl[k], 31);
      }
      unsigned long long* wt64 = reinterpret_cast<unsigned long long*>(scan_sm);
      __syncthreads();                                                        // scan_sm is reused tile after tile
      if (lane == 0) wt64[warp] = wtotal;
      __syncthreads();
      unsigned long long wbase = 0, ttotal = 0;
      #pragma unroll
      for (int w = 0; w < BLOCK_THREADS / 32; w++) { const unsigned long long x = wt64[w]; wbase += w < warp ? x : 0ULL; ttotal += x; }
      unsigned long long o = out_base + wbase;
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        const uint32_t tk = __shfl_sync(0xffffffffu, incl[k], 31);            // (a slot total beyond 2^32 would need 64-bit scans: 2^27 matches per lane)
        const uint32_t excl = incl[k] - m[k];
        for (uint32_t p0 = 0; p0 < tk; p0 += 32) {
          const uint32_t p = p0 + lane;
          int owner = 0;
          #pragma unroll
          for (int step = 16; step > 0; step >>= 1) { const uint32_t v = __shfl_sync(0xffffffffu, incl[k], owner + step - 1); if (v <= p) owner += step; }
          owner = owner > 31 ? 31 : owner;
          const uint32_t s0 = __shfl_sync(0xffffffffu, start[k], owner), e0 = __shfl_sync(0xffffffffu, excl, owner), pr = __shfl_sync(0xffffffffu, prow[k], owner);
          if (p < tk) {
            if (outR) st_stream_u32(outR + o + p, rows[s0 + (p - e0)], pol_s);
            st_stream_u32(outS + o + p, pr, pol_s);
          }
        }
        o += tk;
      }
      out_base += ttotal;
    }
  }
  __syncthreads();                                                            // shared scratch is reused by the next chunk
  }
}

// The write pass reads the table header and the scratch counters back (one sync) and launches the ONE kernel of the path the count
// pass took: grouped / radix by the table's layout, hit lists by the device-side flag, range by all_present, else the match cache.
cudaError_t write_pairs(const void* S, int64_t nS, int key_bytes, const void* table, const void* scratch,
                        int32_t* outR, int32_t* outS, const uint32_t* probe_payload, uint32_t probe_row_base, cudaStream_t stream) {
  ScratchView sv = scratch_view(const_cast<void*>(scratch), nS, key_bytes);
  if (nS == 0) return cudaSuccess;
  const TableHeader* hdr = reinterpret_cast<const TableHeader*>(table);
  const char* body = reinterpret_cast<const char*>(table) + HEADER_BYTES;
  TableHeader h; unsigned long long ctr[SCRATCH_COUNTERS];
  {
    char* pin = reinterpret_cast<char*>(pinned_block());
    if (!pin) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaMemcpyAsync(pin, table, sizeof(TableHeader), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pin + HEADER_BYTES, sv.counters, sizeof(ctr), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    memcpy(&h, pin, sizeof(h)); memcpy(ctr, pin + HEADER_BYTES, sizeof(ctr));
  }
  if (h.magic != HJ_MAGIC || (int)h.key_bytes != key_bytes) return cudaErrorInvalidValue;
  if (ctr[CTR_TOTAL] == 0) return cudaSuccess;                      // nothing to emit (join_v1.mlir:600-601 skips the probe too)
  { cudaError_t e = cudaMemsetAsync(sv.counters + CTR_TICKET_GROUP_W, 0, 8, stream);                 // the write pass may be repeated after one count
    if (e == cudaSuccess) e = cudaMemsetAsync(sv.counters + CTR_TICKET_RADIX_W, 0, 8, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(sv.counters + CTR_TICKET_RADIX_E, 0, 8, stream);
    if (e != cudaSuccess) return e; }
  const unsigned long long* sparse_flag = sv.counters + CTR_SPARSE;
  if (h.mode == MODE_RADIX) {
    if (ctr[CTR_CARRIED]) { probe_payload = nullptr; probe_row_base = 0; }           // the partitioned copy already holds the probe row ids
    return radix_write(nS, key_bytes, h, body, sv.radix, sv.chunk_offsets, sv.counters + CTR_TICKET_RADIX_W, sv.counters + CTR_TICKET_RADIX_E, sv.mcache, sv.counters + CTR_RADIX_MULTI,
                       outR, outS, ctr[CTR_CARRIED] != 0, probe_payload, probe_row_base, ctr[CTR_SEMI] != 0, stream);
  }
  if (ctr[CTR_SLICED]) {                                             // the match cache / hit lists refer to the slice-ordered copy
    const SliceArea sa = slice_area(sv.radix, nS, key_bytes);
    S = sa.keys;
    if (ctr[CTR_CARRIED]) probe_payload = sa.rows;                   // the copy carries the probe row ids themselves
    else {                                                           // ... or original indices: translate once per write (the count pass was given no ids)
      k_translate_rows<<<(unsigned)std::min<int64_t>(148 * 16, (nS + BLOCK_THREADS - 1) / BLOCK_THREADS), BLOCK_THREADS, 0, stream>>>(sa.rows, nS, probe_payload, probe_row_base, sa.rows2);
      probe_payload = sa.rows2;
    }
    probe_row_base = 0;
  }
  const bool vec = (reinterpret_cast<uintptr_t>(S) & 15) == 0;
  const bool grouped = h.mode == MODE_GROUP, lists = !grouped && ctr[CTR_SPARSE] != 0, by_range = !grouped && !lists && h.mode == MODE_DENSE && h.all_present;
  if (lists) {
    k_write_sparse<<<(unsigned)std::min<int64_t>(PERSIST_GRID, (sv.nchunks + BLOCK_THREADS / 32 - 1) / (BLOCK_THREADS / 32)), BLOCK_THREADS, 0, stream>>>(
        hdr, sv.hit_list, sv.warp_counts, sv.chunk_offsets, sv.nchunks, chunk_keys(key_bytes), outR, outS, probe_payload, probe_row_base, sparse_flag);
    return cudaGetLastError();
  }
#define HJ_LAUNCH_WRITE(K, V) \
  if (grouped) k_write<K, V, true><<<resident_grid(k_write<K, V, true>, sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.mcache, sv.run_start, sv.chunk_offsets, sv.nchunks, sv.counters + CTR_TICKET_GROUP_W, outR, outS, probe_payload, probe_row_base, sparse_flag); \
  else if (by_range) k_write_range<K, V><<<range_grid(sv.nchunks), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.chunk_offsets, sv.mcache, sv.nchunks, outR, outS, probe_payload, probe_row_base, sparse_flag); \
  else k_write<K, V, false><<<dense_grid(k_write<K, V, false>, sv.nchunks, g_write_waves), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, sv.mcache, sv.run_start, sv.chunk_offsets, sv.nchunks, sv.counters + CTR_TICKET_GROUP_W, outR, outS, probe_payload, probe_row_base, sparse_flag);
  if (key_bytes == 4) { if (vec) { HJ_LAUNCH_WRITE(int32_t, true) } else { HJ_LAUNCH_WRITE(int32_t, false) } }
  else                { if (vec) { HJ_LAUNCH_WRITE(int64_t, true) } else { HJ_LAUNCH_WRITE(int64_t, false) } }
#undef HJ_LAUNCH_WRITE
  return cudaGetLastError();
}

// =========================================================================================================
// K2+K3+K4 fused: single-pass probe for callers that can bound the result size up front (e.g. capacity = |S| for a unique
// build). One kernel looks up a tile, publishes its match count, obtains its output offset by DECOUPLED LOOK-BACK over the
// tiles before it (tiles are taken by ticket, so every tile a CTA waits for is already running) and streams its pairs
// straight into the result columns: no match cache, no second pass over the probe relation. The reference's call sequence
// (countRows -> allocate -> probeRelation, join_v1.mlir:591,604-605) cannot use it; hjJoinFused can.
// (Measured and rejected in round 2: publishing a tile's aggregate from the range test alone, before its lookups return, with the next
// ticket fetched ahead: config 2 2.44 -> 3.68 ms.)
// =========================================================================================================
constexpr unsigned long long LB_AGG = 1ULL << 62, LB_INCL = 2ULL << 62, LB_MASK = (1ULL << 62) - 1;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

template <typename K, bool VEC, uint32_t MODE>
__global__ void __launch_bounds__(BLOCK_THREADS) k_join_fused(const K* __restrict__ S, int64_t nS, const char* __restrict__ body, const TableHeader* __restrict__ hdr,
                                                              unsigned long long* __restrict__ tile_state, unsigned long long* tickets, int64_t ntiles,
                                                              int32_t* __restrict__ outR, int32_t* __restrict__ outS, unsigned long long capacity,
                                                              const uint32_t* __restrict__ probe_payload, uint32_t probe_row_base,
                                                              unsigned long long* __restrict__ total_out) {
  using T = KeyTraits<K>;
  if (hdr->mode != MODE) return;
  constexpr int KPV = T::KEYS_PER_VEC, KPT = VECS_PER_THREAD * KPV, TILE = BLOCK_THREADS * KPT, WARPS = BLOCK_THREADS / 32;
  __shared__ long long ticket;
  __shared__ uint32_t wt[WARPS];
  __shared__ unsigned long long base_sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t n_pairs = hdr->n_pairs;
  const long long kmin = hdr->kmin;
  const unsigned long long drange = hdr->dense_range;
  const uint64_t pol_s = policy_evict_first(), pol_t = policy_evict_last();

  for (long long tile = next_ticket(tickets, &ticket); tile < ntiles; tile = next_ticket(tickets, &ticket)) {
    const int64_t tile_base = tile * TILE;
    K key[KPT];
    #pragma unroll
    for (int v = 0; v < VECS_PER_THREAD; v++) load_vec_keys<K, VEC>(S, nS, tile_base + ((int64_t)v * BLOCK_THREADS + threadIdx.x) * KPV, pol_s, &key[v * KPV]);
    uint32_t m[KPT];
    if constexpr (MODE == MODE_DENSE) {
      #pragma unroll
      for (int k = 0; k < KPT; k++) {
        const unsigned long long off = (unsigned long long)((long long)key[k] - kmin);
        m[k] = (off < drange && elem_index<KPV>(tile_base, k) < nS) ? ld_keep_u32(reinterpret_cast<const uint32_t*>(body) + off, pol_t) : ROW_NONE;
      }
    } else {
      #pragma unroll
      for (int v = 0; v < VECS_PER_THREAD; v++) {
        Bucket b[KPV];
        #pragma unroll
        for (int e = 0; e < KPV; e++) b[e] = ld_bucket(home_bucket<K>(body, n_pairs, key[v * KPV + e]));
        #pragma unroll
        for (int e = 0; e < KPV; e++) {
          const int k = v * KPV + e;
          m[k] = elem_index<KPV>(tile_base, k) < nS ? finish_probe_unique<K>(body, n_pairs, key[k], b[e]) : ROW_NONE;
        }
      }
    }
    unsigned mask[KPT];
    uint32_t wtot = 0;
    #pragma unroll
    for (int k = 0; k < KPT; k++) { mask[k] = __ballot_sync(0xffffffffu, m[k] != ROW_NONE); wtot += __popc(mask[k]); }
    if (lane == 0) wt[warp] = wtot;
    __syncthreads();
    uint32_t wbase = 0, ttot = 0;
    #pragma unroll
    for (int w = 0; w < WARPS; w++) { const uint32_t x = wt[w]; wbase += w < warp ? x : 0u; ttot += x; }

    // decoupled look-back, one warp: publish this tile's aggregate, then sum the tiles before it until an inclusive prefix shows up
    if (warp == 0) {
      if (lane == 0) st_release_u64(tile_state + tile, (tile == 0 ? LB_INCL : LB_AGG) | ttot);
      unsigned long long prefix = 0;
      long long j = tile - 1;
      while (j >= 0) {
        const long long idx = j - lane;
        unsigned long long st = idx >= 0 ? ld_acquire_u64(tile_state + idx) : LB_INCL;
        while (__any_sync(0xffffffffu, (st >> 62) == 0)) { if ((st >> 62) == 0) st = ld_acquire_u64(tile_state + idx); }
        const unsigned incl = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = incl ? __ffs(incl) - 1 : 32;
        unsigned long long v = lane <= first ? (st & LB_MASK) : 0ULL;
        v = warp_reduce_sum(v);
        prefix += v;
        if (first < 32) break;
        j -= 32;
      }
      if (lane == 0) {
        if (tile > 0) st_release_u64(tile_state + tile, LB_INCL | (prefix + ttot));
        base_sm = prefix;
        if (tile == ntiles - 1) *total_out = prefix + ttot;
      }
    }
    __syncthreads();
    unsigned long long o = base_sm + wbase;
    const unsigned lt = (1u << lane) - 1u;
    #pragma unroll
    for (int k = 0; k < KPT; k++) {
      if (m[k] != ROW_NONE) {
        const unsigned long long dst = o + __popc(mask[k] & lt);
        if (dst < capacity) {
          const int64_t j = elem_index<KPV>(tile_base, k);
          st_stream_u32(outR + dst, m[k], pol_s);
          st_stream_u32(outS + dst, probe_payload ? probe_payload[j] : probe_row_base + (uint32_t)j, pol_s);
        }
      }
      o += __popc(mask[k]);
    }
  }
}

// Returns through total_out (device, u64): pairs found (pairs beyond `capacity` are counted but not written). *unsupported is set
// when the table is in the grouped layout (duplicate build keys): the caller falls back to count + write.
cudaError_t join_fused_async(const void* S, int64_t nS, int key_bytes, const void* table, void* scratch, int32_t* outR, int32_t* outS, int64_t capacity,
                             const uint32_t* probe_payload, uint32_t probe_row_base, cudaStream_t stream) {
  ScratchView sv = scratch_view(scratch, nS, key_bytes);
  const TableHeader* hdr = reinterpret_cast<const TableHeader*>(table);
  const char* body = reinterpret_cast<const char*>(table) + HEADER_BYTES;
  const int64_t ntiles = (nS + tile_keys(key_bytes) - 1) / tile_keys(key_bytes);
  unsigned long long* tile_state = reinterpret_cast<unsigned long long*>(sv.mcache);      // the match cache is not needed: reuse it (8 B per tile)
  unsigned long long* total = sv.counters + CTR_TOTAL;                                    // same slot the two-phase path reports in
  cudaError_t e = cudaMemsetAsync(sv.counters, 0, SCRATCH_COUNTERS * sizeof(unsigned long long), stream);
  if (e == cudaSuccess && ntiles > 0) e = cudaMemsetAsync(tile_state, 0, (size_t)ntiles * 8, stream);
  if (e != cudaSuccess || ntiles == 0) return e;
  const bool vec = (reinterpret_cast<uintptr_t>(S) & 15) == 0;
#define HJ_LAUNCH_FUSED(K, V) \
  k_join_fused<K, V, MODE_DENSE><<<resident_grid(k_join_fused<K, V, MODE_DENSE>, ntiles), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, tile_state, sv.counters, ntiles, outR, outS, (unsigned long long)capacity, probe_payload, probe_row_base, total); \
  k_join_fused<K, V, MODE_HASH><<<resident_grid(k_join_fused<K, V, MODE_HASH>, ntiles), BLOCK_THREADS, 0, stream>>>((const K*)S, nS, body, hdr, tile_state, sv.counters + CTR_TICKET_GROUP, ntiles, outR, outS, (unsigned long long)capacity, probe_payload, probe_row_base, total);
  if (key_bytes == 4) { if (vec) { HJ_LAUNCH_FUSED(int32_t, true) } else { HJ_LAUNCH_FUSED(int32_t, false) } }
  else                { if (vec) { HJ_LAUNCH_FUSED(int64_t, true) } else { HJ_LAUNCH_FUSED(int64_t, false) } }
#undef HJ_LAUNCH_FUSED
  return cudaGetLastError();
}

// =========================================================================================================
// K6  digest + generators
// =========================================================================================================
__global__ void __launch_bounds__(BLOCK_THREADS) k_pair_digest(const int32_t* __restrict__ outR, const int32_t* __restrict__ outS, int64_t n,
                                                               unsigned long long* __restrict__ out2) {
  __shared__ unsigned long long rs[32], rx[32];
  unsigned long long s = 0, x = 0;
  for (int64_t i = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK_THREADS) {
    const uint64_t m = mix64(((uint64_t)(uint32_t)outR[i] << 32) | (uint32_t)outS[i]);
    s += m; x ^= m;
  }
  #pragma unroll
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); x ^= __shfl_xor_sync(0xffffffffu, x, d); }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rx[threadIdx.x >> 5] = x; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BLOCK_THREADS / 32; w++) { s += rs[w]; x ^= rx[w]; }
    atomicAdd(&out2[0], s); atomicXor(&out2[1], x);
  }
}
cudaError_t pair_digest(const int32_t* outR, const int32_t* outS, int64_t n, unsigned long long* out2, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out2, 0, 16, stream);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    int64_t blocks = (n + BLOCK_THREADS - 1) / BLOCK_THREADS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_pair_digest<<<(unsigned)blocks, BLOCK_THREADS, 0, stream>>>(outR, outS, n, out2);
  }
  return cudaGetLastError();
}

// ---- generators: integer-only, bit-identical to oracle/oracle_join.c (gen_one) ----
__device__ __forceinline__ uint64_t rnd64(uint64_t seed, uint64_t i) { return mix64(seed * 0x9E3779B97F4A7C15ULL + mix64(i + 0xD1B54A32D192ED03ULL)); }
__host__ __device__ inline int half_bits(uint64_t n) { int b = 1; while (b < 64 && ((uint64_t)1 << b) < n) b++; return (b + 1) / 2; }
__device__ __forceinline__ uint64_t feistel_fwd(uint64_t x, int hb, uint64_t seed) {
  const uint64_t mask = ((uint64_t)1 << hb) - 1;
  uint64_t l = x >> hb, r = x & mask;
  #pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint64_t f = mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(k + 1));
    const uint64_t t = l ^ (mix64(r ^ f) & mask); l = r; r = t;
  }
  return (l << hb) | r;
}
__device__ __forceinline__ uint64_t perm(uint64_t i, uint64_t n, int hb, uint64_t seed) {
  uint64_t x = i;
  do { x = feistel_fwd(x, hb, seed); } while (x >= n);
  return x;
}
struct ZipfTable { unsigned long long t[257]; };
__device__ __forceinline__ uint64_t zipf_rank(uint64_t r, int log2D, const unsigned long long* t) {
  const uint32_t a = (uint32_t)(((r >> 32) * (uint64_t)log2D) >> 32);
  const uint32_t f = (uint32_t)r & 0xFFFF, hi = f >> 8, lo = f & 0xFF;
  const uint64_t m = t[hi] + (((t[hi + 1] - t[hi]) * lo) >> 8);
  return m >> (62 - a);
}

template <typename K>
__global__ void __launch_bounds__(BLOCK_THREADS) k_generate(K* __restrict__ out, int64_t n, int kind, uint64_t seed, int64_t lo, uint64_t domain, uint32_t p16,
                                                            uint64_t key_mul, int64_t index_base, uint64_t n_total, int hb_domain, int hb_total, int log2D,
                                                            const ZipfTable zt, const uint32_t* __restrict__ at) {
  __shared__ unsigned long long zsm[257];
  for (int t = threadIdx.x; t < 257; t += BLOCK_THREADS) zsm[t] = zt.t[t];
  __syncthreads();
  for (int64_t li = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; li < n; li += (int64_t)gridDim.x * BLOCK_THREADS) {
    const uint64_t i = at ? (uint64_t)at[li] : (uint64_t)(index_base + li);          // at: the key of row at[li] instead of row index_base + li
    uint64_t v;
    switch (kind) {
      case 0: v = i; break;
      case 1: v = perm(i, domain, hb_domain, seed); break;
      case 2: v = __umul64hi(rnd64(seed, i), domain); break;
      case 3: { const uint64_t r = rnd64(seed, i); const uint64_t u = __umul64hi(rnd64(seed ^ 0xA5A5A5A5ULL, i), domain);
                v = ((r & 0xFFFF) < p16) ? u : domain + u; break; }
      case 4: v = perm(i, n_total, hb_total, seed) % domain; break;
      case 5: v = perm(zipf_rank(rnd64(seed, i), log2D, zsm) - 1, domain, hb_domain, seed ^ 0x5EEDULL); break;
      default: v = 0;
    }
    int64_t val = lo + (int64_t)v;
    if (key_mul) val = sizeof(K) == 8 ? (int64_t)((uint64_t)val * key_mul) : (int64_t)(int32_t)((uint32_t)val * (uint32_t)key_mul);
    out[li] = (K)val;
  }
}

cudaError_t generate_keys(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                          uint32_t p16, uint64_t key_mul, int64_t index_base, cudaStream_t stream);   // (n_total == n + index_base unless set)

static void make_zipf_table(ZipfTable& z) {
  z.t[0] = 1ULL << 62;
  for (int i = 1; i <= 256; i++) z.t[i] = (unsigned long long)(((unsigned __int128)z.t[i - 1] * 0x8058D7D2D5E5F6B1ULL) >> 63);
  z.t[256] = 1ULL << 63;
}

cudaError_t generate_keys_total(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                                uint32_t p16, uint64_t key_mul, int64_t index_base, uint64_t n_total, const uint32_t* at, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (domain == 0) domain = 1;
  ZipfTable zt; make_zipf_table(zt);
  int log2D = 0; while ((2ULL << log2D) <= domain) log2D++;
  int64_t blocks = (n + BLOCK_THREADS - 1) / BLOCK_THREADS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int hbd = half_bits(domain), hbt = half_bits(n_total ? n_total : 1);
  if (key_bytes == 4) k_generate<int32_t><<<(unsigned)blocks, BLOCK_THREADS, 0, stream>>>((int32_t*)out, n, kind, seed, lo, domain, p16, key_mul, index_base, n_total, hbd, hbt, log2D, zt, at);
  else                k_generate<int64_t><<<(unsigned)blocks, BLOCK_THREADS, 0, stream>>>((int64_t*)out, n, kind, seed, lo, domain, p16, key_mul, index_base, n_total, hbd, hbt, log2D, zt, at);
  return cudaGetLastError();
}
cudaError_t generate_keys(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                          uint32_t p16, uint64_t key_mul, int64_t index_base, cudaStream_t stream) {
  return generate_keys_total(out, n, key_bytes, kind, seed, lo, domain, p16, key_mul, index_base, (uint64_t)(index_base + n), nullptr, stream);
}

}  // namespace hj
