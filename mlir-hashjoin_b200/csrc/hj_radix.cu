// hj_radix.cu — K7: radix join for build relations whose hash table would not stay in L2 (> 48 MB of buckets).
//
// Replaces, for those sizes, the global open-addressing table (K1/K2/K4 in hj_kernels.cu; the reference's chained table,
// join_v1.mlir:25-39,213-277) by the plan the reference lists as left out (projectDescription.md:24, "partitioned hash-join"):
//   build   both passes of K5 (hj_partition.cu) split the build relation into 2^(b1+b2) partitions of <= ~4 096 (key, row id) tuples on
//           the top bits of radix_hash(key); the partitioned copy + its offsets ARE the table (no clear, no global atomics);
//   count   the probe relation is partitioned the same way; a work item = (partition, <= 16 384 probe tuples of it). A CTA takes items
//           by ticket, builds the partition's table in SHARED memory — 4 096 buckets in CSR form: one ATOMS.ADD per tuple ranks it inside
//           its bucket, a block scan turns bucket counts into a directory (first entry, entries), the tuples land bucket-sorted; no sentinel, so every key value
//           stays legal, and no probe-sequence clustering, so duplicate keys cost exactly their multiplicity — streams its probe
//           tuples through it and writes the item's match count; K3 (the same scan as the other layouts) turns item counts into
//           offsets and the total;
//           the count pass also leaves a MATCH CACHE word per partitioned probe tuple (position of its matched build tuple, or none)
//           and flags the items where some probe tuple matched more than once;
//   write   items whose probe tuples match at most once (unique build keys, semi-joins) are a pure stream: k_rj_emit reads cache words
//           and row ids, ranks the hits of 32 tuples with a ballot and claims a run of the item's output range with one shared-memory
//           atomic — no table, no keys. The flagged items are joined again (k_rj_join<write>: table rebuilt with row ids, matches of a
//           warp's 32 tuples pre-counted so that one atomic claims their run). Order is free (shared.cpp:168-171).
// Work items are handed out by ticket two ahead, so that a CTA pulls its next item's tuples into L2 while it joins the current one.
// Duplicate build keys need nothing special (equal keys share a bucket and every entry of it is compared); a partition larger than the
// table (heavy skew) is joined in rounds of RJ_CAP build tuples; a hot probe key only makes more items.
#include <algorithm>
#include <cstdio>
#include "hj_common.cuh"
#include "hj_kernels.cuh"

namespace hj {

constexpr int RJ_THREADS = 512;
constexpr int RJ_BUCKETS = 4096;        // shared-memory table: buckets (CSR starts) ...
constexpr int RJ_CAP = 4864;            // ... over at most this many build tuples per round: 72 KB (i64) / 53 KB (i32) per CTA, 3 / 4 CTAs per SM
constexpr int RJ_R_ITEMS = (RJ_CAP + RJ_THREADS - 1) / RJ_THREADS;
constexpr int RJ_TARGET = 4096;         // partitions are sized so that the AVERAGE build partition is at most this (sigma = 64 for uniform hashes)
constexpr int RJ_ITEM_ROWS = 16384;     // probe tuples per work item
constexpr int RJ_MAX_BITS = 16;         // two passes of <= 8 bits

static inline int64_t r256(int64_t x) { return (x + 255) / 256 * 256; }

void radix_bits(int64_t n_build, int* bits1, int* bits2) {
  int bits = 2;
  while (bits < RJ_MAX_BITS && ((int64_t)RJ_TARGET << bits) < n_build) bits++;
  *bits1 = (bits + 1) / 2; *bits2 = bits - *bits1;
}
int64_t radix_max_items(int64_t n_probe) { return (n_probe + RJ_ITEM_ROWS - 1) / RJ_ITEM_ROWS + ((int64_t)1 << RJ_MAX_BITS) + 1; }

// ---- table body in the radix layout: [keys][row ids][offsets u32 x (parts + 1)][pass-1 keys][pass-1 row ids][partition workspace] ----
struct RadixArea { char* keys; uint32_t* rows; uint32_t* offsets; char* tmp_keys; uint32_t* tmp_rows; void* ws; int64_t ws_bytes; };
static int64_t radix_area_bytes(int64_t n, int key_bytes, int bits1, int bits2) {
  return 2 * (r256(n * key_bytes) + r256(n * 4)) + r256((((int64_t)1 << (bits1 + bits2)) + 1) * 4) + radix_partition2_workspace_bytes(n, bits1, bits2);
}
static RadixArea radix_area(char* base, int64_t n, int key_bytes, int bits1, int bits2) {
  RadixArea a;
  a.keys = base; base += r256(n * key_bytes);
  a.rows = reinterpret_cast<uint32_t*>(base); base += r256(n * 4);
  a.offsets = reinterpret_cast<uint32_t*>(base); base += r256((((int64_t)1 << (bits1 + bits2)) + 1) * 4);
  a.tmp_keys = base; base += r256(n * key_bytes);
  a.tmp_rows = reinterpret_cast<uint32_t*>(base); base += r256(n * 4);
  a.ws = base; a.ws_bytes = radix_partition2_workspace_bytes(n, bits1, bits2);
  return a;
}
int64_t radix_table_bytes(int64_t n_build, int key_bytes) {
  int b1, b2; radix_bits(n_build, &b1, &b2);
  return radix_area_bytes(n_build, key_bytes, b1, b2);
}
// probe side: the partition count is the TABLE's, unknown when the scratch is sized: room for the most partitions there can be
struct RadixScratch { RadixArea a; unsigned long long* item_start; RjItem* items; unsigned char* item_multi; int64_t max_items; };
static int64_t radix_items_bytes(int64_t n_probe) {
  return r256((((int64_t)1 << RJ_MAX_BITS) + 1 + 264) * 8) + r256(radix_max_items(n_probe) * (int64_t)sizeof(RjItem)) + r256(radix_max_items(n_probe));
}
int64_t radix_scratch_bytes(int64_t n_probe, int key_bytes) { return radix_area_bytes(n_probe, key_bytes, 8, 8) + radix_items_bytes(n_probe); }
static RadixScratch radix_scratch(char* base, int64_t n_probe, int key_bytes, int bits1, int bits2) {
  RadixScratch s;
  s.max_items = radix_max_items(n_probe);
  s.item_start = reinterpret_cast<unsigned long long*>(base);
  s.items = reinterpret_cast<RjItem*>(base + r256((((int64_t)1 << RJ_MAX_BITS) + 1 + 264) * 8));
  s.item_multi = reinterpret_cast<unsigned char*>(s.items) + r256(s.max_items * (int64_t)sizeof(RjItem));
  s.a = radix_area(base + radix_items_bytes(n_probe), n_probe, key_bytes, bits1, bits2);
  return s;
}

// The slice-ordered inline layout (hj_kernels.cu) keeps its partitioned copy of the probe relation in the same scratch area: it needs
// one copy of (key, row id), 257 offsets, one more row-id array and a one-pass workspace — all smaller than what the radix layout has there.
SliceArea slice_area(char* base, int64_t n, int key_bytes) {
  SliceArea a;
  a.keys = base; base += r256(n * key_bytes);
  a.rows = reinterpret_cast<uint32_t*>(base); base += r256(n * 4);
  a.offsets = reinterpret_cast<uint32_t*>(base); base += r256(257 * 4);
  a.rows2 = reinterpret_cast<uint32_t*>(base); base += r256(n * 4);
  a.ws = base; a.ws_bytes = slice_partition_workspace_bytes(n, 8);
  return a;
}

// ---------------------------------------------------------------------------------------------------------
// build: partition the build relation, record the layout in the header
// ---------------------------------------------------------------------------------------------------------
__global__ void k_rj_header(TableHeader* hdr, uint32_t bits1, uint32_t bits2, unsigned long long keys_off, unsigned long long rows_off, unsigned long long offs_off) {
  hdr->mode = MODE_RADIX; hdr->rj_bits1 = bits1; hdr->rj_bits2 = bits2; hdr->rj_keys_off = keys_off; hdr->rj_rows_off = rows_off; hdr->rj_offs_off = offs_off;
}

cudaError_t radix_build(const void* R, int64_t nR, int key_bytes, const uint32_t* payload, uint32_t row_base, TableHeader* hdr, char* body, int64_t body_bytes,
                        cudaStream_t stream) {
  int b1, b2; radix_bits(nR, &b1, &b2);
  if (body_bytes < radix_area_bytes(nR, key_bytes, b1, b2)) return cudaErrorInvalidValue;
  const RadixArea a = radix_area(body, nR, key_bytes, b1, b2);
  cudaError_t e = radix_partition2(R, payload, row_base, nR, key_bytes, b1, b2, a.tmp_keys, a.tmp_rows, a.keys, a.rows, a.offsets, a.ws, a.ws_bytes, stream);
  if (e != cudaSuccess) return e;
  k_rj_header<<<1, 1, 0, stream>>>(hdr, (uint32_t)b1, (uint32_t)b2, (unsigned long long)(a.keys - body), (unsigned long long)(reinterpret_cast<char*>(a.rows) - body),
                                   (unsigned long long)(reinterpret_cast<char*>(a.offsets) - body));
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// work items
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rj_item_counts(const uint32_t* __restrict__ offR, const uint32_t* __restrict__ offS, uint32_t n_parts, unsigned long long* __restrict__ item_start) {
  const uint32_t q = blockIdx.x * 256 + threadIdx.x;
  if (q >= n_parts) return;
  const uint32_t ns = offS[q + 1] - offS[q], nr = offR[q + 1] - offR[q];
  item_start[q] = (ns && nr) ? (ns + RJ_ITEM_ROWS - 1) / RJ_ITEM_ROWS : 0u;      // a partition without build tuples matches nothing: no items
}
__global__ void __launch_bounds__(256) k_rj_items(const uint32_t* __restrict__ offR, const uint32_t* __restrict__ offS, uint32_t n_parts,
                                                  const unsigned long long* __restrict__ item_start, RjItem* __restrict__ items) {
  const uint32_t q = blockIdx.x * 256 + threadIdx.x;
  if (q >= n_parts) return;
  const unsigned long long first = item_start[q], cnt = item_start[q + 1] - first;
  const uint32_t s0 = offS[q], s1 = offS[q + 1], r0 = offR[q], r1 = offR[q + 1];
  for (unsigned long long j = 0; j < cnt; j++) {
    const uint32_t a = s0 + (uint32_t)j * RJ_ITEM_ROWS;
    items[first + j] = RjItem{a, s1 - a < (uint32_t)RJ_ITEM_ROWS ? s1 : a + RJ_ITEM_ROWS, r0, r1};
  }
}

// ---------------------------------------------------------------------------------------------------------
// the join kernel: WRITE = false counts the matches of every item, WRITE = true emits them at the item's offset
// ---------------------------------------------------------------------------------------------------------
// bucket directory: one word per bucket, first entry << 16 | entries (both < 2^13), so a probe needs ONE shared-memory load to know its
// candidates. The scan reads 8 consecutive buckets per thread: one pad word per 32 keeps those accesses (and everyone else's) conflict-free.
__device__ __forceinline__ uint32_t spad(uint32_t b) { return b + (b >> 5); }
constexpr int RJ_DIR_WORDS = RJ_BUCKETS + RJ_BUCKETS / 32 + 1;
constexpr int RJ_BUCKET_SHIFT = 32 - 12;                 // bucket = top 12 bits of bucket_hash
static_assert(RJ_BUCKETS == 1 << 12 && RJ_CAP < (1 << 13), "bucket shift / directory packing");
template <typename K> struct RjSmem { K ckey[RJ_CAP]; uint32_t crow[RJ_CAP]; uint32_t dir[RJ_DIR_WORDS]; };

template <typename K>
__device__ __forceinline__ void rj_build_round(RjSmem<K>& sm, const K* __restrict__ Rk, const uint32_t* __restrict__ Rr, uint32_t r0, uint32_t nr, bool with_rows,
                                               uint64_t pol, uint32_t* scan_sm) {
  constexpr int PER = RJ_BUCKETS / RJ_THREADS;                       // directory words per thread in the scan
  for (int i = threadIdx.x; i < RJ_DIR_WORDS; i += RJ_THREADS) sm.dir[i] = 0;
  __syncthreads();
  uint32_t br[RJ_R_ITEMS];                                           // bucket << 16 | rank inside the bucket
  // All loads first, then the atomics: with load -> hash -> atomic per tuple the compiler keeps the loads behind the atomics and every one
  // of them exposes its own latency (ncu, 2^28 x 2^28 i64: long-scoreboard stalls at each of the unrolled hashes, 24 % of the kernel).
  #pragma unroll
  for (int u = 0; u < RJ_R_ITEMS; u++) {
    const uint32_t i = u * RJ_THREADS + threadIdx.x;
    br[u] = i < nr ? bucket_hash<K>(Rk[r0 + i]) >> RJ_BUCKET_SHIFT : 0xFFFFFFFFu;   // default cache policy: the second look below must hit L2
  }
  #pragma unroll
  for (int u = 0; u < RJ_R_ITEMS; u++)
    if (br[u] != 0xFFFFFFFFu) br[u] = (br[u] << 16) | atomicAdd(&sm.dir[spad(br[u])], 1u);   // the count sits in the low half: the returned value is the rank
  __syncthreads();
  {                                                                   // counts -> first << 16 | count: thread t owns buckets [t * PER, (t + 1) * PER)
    uint32_t v[PER], sum = 0;
    #pragma unroll
    for (int i = 0; i < PER; i++) { v[i] = sm.dir[spad(threadIdx.x * PER + i)]; sum += v[i]; }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, scan_sm, &total);
    #pragma unroll
    for (int i = 0; i < PER; i++) { sm.dir[spad(threadIdx.x * PER + i)] = (run << 16) | v[i]; run += v[i]; }
  }
  __syncthreads();
  #pragma unroll
  for (int u = 0; u < RJ_R_ITEMS; u++) {
    const uint32_t i = u * RJ_THREADS + threadIdx.x;
    if (br[u] != 0xFFFFFFFFu) {                                       // second look at the tuple: an L1 / L2 hit (the partition was read a moment ago)
      const uint32_t pos = (sm.dir[spad(br[u] >> 16)] >> 16) + (br[u] & 0xFFFFu);
      sm.ckey[pos] = ld_stream<K>(Rk + r0 + i, pol);
      sm.crow[pos] = with_rows ? ld_stream<uint32_t>(Rr + r0 + i, pol) : i;   // counting keeps the tuple's position instead: it goes into the match cache
    }
  }
  __syncthreads();
}

template <typename K, bool WRITE>
__global__ void __launch_bounds__(RJ_THREADS, 3) k_rj_join(const K* __restrict__ Rk, const uint32_t* __restrict__ Rr, const K* __restrict__ Sk, const uint32_t* __restrict__ Sr,
                                                           const RjItem* __restrict__ items, const unsigned long long* __restrict__ n_items_ptr, unsigned long long* tickets,
                                                           unsigned long long* __restrict__ item_totals,       // count: out (matches per item); write: in (exclusive offsets)
                                                           int32_t* __restrict__ outR, int32_t* __restrict__ outS,
                                                           const uint32_t* __restrict__ probe_payload, uint32_t probe_row_base, int carried_rows, int semi,
                                                           uint32_t* __restrict__ mcache,                       // count: out, per partitioned probe tuple: matched build position | NONE
                                                           unsigned char* __restrict__ item_multi,              // count: out / write: in — items the match cache cannot describe
                                                           unsigned long long* __restrict__ n_multi) {
  extern __shared__ __align__(16) unsigned char rj_raw[];
  RjSmem<K>& sm = *reinterpret_cast<RjSmem<K>*>(rj_raw);
  __shared__ TicketQueue2 tq;
  __shared__ unsigned long long acc_cnt[2];                        // the item's match count and multi flag, double-buffered over the items
  __shared__ uint32_t acc_multi[2];
  __shared__ uint32_t scan_sm[33];
  __shared__ uint32_t cursor;
  if (threadIdx.x < 2) { acc_cnt[threadIdx.x] = 0; acc_multi[threadIdx.x] = 0; }
  const long long n_items = (long long)*n_items_ptr;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const uint64_t pol = policy_evict_first();
  if (WRITE && *n_multi == 0) return;                             // every item was described by the match cache: k_rj_emit writes them all
  long long item, next;
  ticket2_first(tickets, &tq, item, next);
  for (uint32_t it = 0; item < n_items; it++) {
    const long long pending = ticket_prefetch(tickets);
    if (WRITE && !item_multi[item]) { ticket2_advance(&tq, it, pending, item, next); continue; }      // uniform: k_rj_emit has this item
    const RjItem w = items[item];
    // (count pass only: the write join exists for items with several matches per probe tuple, which are bound by their output — there the
    // extra live registers cost more than the pulls save: 10M x 10M -> 1e9 pairs 2.52 -> 2.73 ms)
    const bool pull_next = !WRITE && next < n_items;
    RjItem wn = w;
    if (pull_next) wn = items[next];                                // in flight during the build below
    unsigned long long cnt = 0;
    unsigned long long obase = 0;
    const bool one_round = w.r1 - w.r0 <= (uint32_t)RJ_CAP;
    bool multi = !one_round && !semi;                               // a key may match in several rounds: no single cache entry can say so
    if (WRITE) { obase = item_totals[item]; if (threadIdx.x == 0) cursor = 0; }
    for (uint32_t r0 = w.r0; r0 < w.r1; r0 += RJ_CAP) {
      const uint32_t nr = w.r1 - r0 < (uint32_t)RJ_CAP ? w.r1 - r0 : (uint32_t)RJ_CAP;
      rj_build_round<K>(sm, Rk, Rr, r0, nr, WRITE, pol, scan_sm);
      if (pull_next && r0 == w.r0) {                                // the next item's tuples into L2 while this one is probed
        const uint32_t nrn = wn.r1 - wn.r0 < (uint32_t)RJ_CAP ? wn.r1 - wn.r0 : (uint32_t)RJ_CAP;
        if (wn.r0 != w.r0) prefetch_l2_range(Rk + wn.r0, nrn * (uint32_t)sizeof(K), RJ_THREADS);
        prefetch_l2_range(Sk + wn.s0, (wn.s1 - wn.s0) * (uint32_t)sizeof(K), RJ_THREADS);
      }
      constexpr int U = 4;
      if (!WRITE) {
        uint32_t c = 0;
        for (uint32_t j0 = w.s0; j0 < w.s1; j0 += RJ_THREADS * U) {
          K key[U];
          #pragma unroll
          for (int u = 0; u < U; u++) { const uint32_t j = j0 + u * RJ_THREADS + threadIdx.x; key[u] = j < w.s1 ? ld_stream<K>(Sk + j, pol) : K(0); }
          #pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t j = j0 + u * RJ_THREADS + threadIdx.x;
            if (j < w.s1) {
              const uint32_t b = bucket_hash<K>(key[u]) >> RJ_BUCKET_SHIFT;
              uint32_t m = 0, first = ROW_NONE;
              const uint32_t de = sm.dir[spad(b)];
              for (uint32_t p = de >> 16, p1 = (de >> 16) + (de & 0xFFFFu); p < p1; p++) {
                if (sm.ckey[p] == key[u]) { if (!m) first = sm.crow[p]; m++; }
              }
              // match cache: the matched build tuple's position in the partitioned build copy (or NONE); with it the write pass of
              // this item is a pure stream (k_rj_emit) — valid while every probe tuple of the item has at most one match
              if (!semi) {
                c += m; multi |= m > 1;
                if (one_round) mcache[j] = m ? r0 + first : ROW_NONE;
              } else {                                                                             // semi-join: a probe tuple counts once, in the first round that matches it
                const uint32_t prev = r0 == w.r0 ? ROW_NONE : mcache[j];                          // this thread's own entry of an earlier round
                if (prev == ROW_NONE) { c += m != 0; if (m || r0 == w.r0) mcache[j] = m ? r0 + first : ROW_NONE; }
              }
            }
          }
        }
        cnt += c;
      } else {
        // warp-synchronous: every warp walks the probe sequences of its 32 tuples in lockstep; the matches of one step are ranked
        // with a ballot and get a contiguous run of the item's output range from ONE shared-memory atomic
        for (uint32_t j0 = w.s0 + (threadIdx.x & ~31u); j0 < w.s1; j0 += RJ_THREADS * U) {        // warp-uniform bounds
          K key[U]; uint32_t srow[U];
          #pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t j = j0 + u * RJ_THREADS + lane;
            key[u] = j < w.s1 ? ld_stream<K>(Sk + j, pol) : K(0);
            srow[u] = j < w.s1 ? ld_stream<uint32_t>(Sr + j, pol) : 0u;
          }
          #pragma unroll
          for (int u = 0; u < U; u++) {
            if (j0 + u * RJ_THREADS >= w.s1) break;                                               // warp-uniform
            bool active = j0 + u * RJ_THREADS + lane < w.s1;
            uint32_t prow = srow[u];
            if (!carried_rows) prow = probe_payload ? (active ? probe_payload[prow] : 0u) : probe_row_base + prow;   // the copy carries original indices
            const uint32_t b = bucket_hash<K>(key[u]) >> RJ_BUCKET_SHIFT;
            const uint32_t de = active ? sm.dir[spad(b)] : 0u;
            uint32_t p = de >> 16;
            const uint32_t p1 = (de >> 16) + (de & 0xFFFFu);
            // first look: how many pairs these 32 tuples make, so that ONE shared-memory atomic claims the warp's run and the lockstep
            // walk below ranks with a ballot and a register cursor only (with an atomic + shuffle per step the dependent latency of every
            // step was on the critical path: config 4's write pass 3.0 ms for 0.95 ms in the count pass)
            uint32_t m = 0;
            for (uint32_t q = p; q < p1; q++) m += sm.ckey[q] == key[u];
            if (semi && m) m = 1;
            const uint32_t total = warp_reduce_sum(m);
            if (total == 0) continue;                                                             // warp-uniform
            uint32_t run = 0;
            if (lane == 0) run = atomicAdd(&cursor, total);
            run = __shfl_sync(0xffffffffu, run, 0);
            if (m == 0) p = p1;                                                                   // nothing to find for this lane
            while (__any_sync(0xffffffffu, p < p1)) {
              bool hit = false; uint32_t brow = 0;
              if (p < p1) { hit = sm.ckey[p] == key[u]; brow = sm.crow[p]; p = (semi && hit) ? p1 : p + 1; }
              const unsigned hm = __ballot_sync(0xffffffffu, hit);
              if (hit) {
                const unsigned long long pos = obase + run + __popc(hm & lt);
                if (outR) outR[pos] = (int32_t)brow;
                outS[pos] = (int32_t)prow;
              }
              run += __popc(hm);
            }
          }
        }
      }
      __syncthreads();                                            // the table is rebuilt by the next round / item
    }
    const long long done = item;
    if (!WRITE) {                                                   // one shared-memory atomic per warp; the barrier of ticket_advance closes the item
      cnt = warp_reduce_sum(cnt);
      const unsigned wm = __ballot_sync(0xffffffffu, multi);
      if (lane == 0) { atomicAdd(&acc_cnt[it & 1], cnt); if (wm) acc_multi[it & 1] = 1; }
    }
    ticket2_advance(&tq, it, pending, item, next);
    if (!WRITE && threadIdx.x == 0) {
      item_totals[done] = acc_cnt[it & 1];
      if (acc_multi[it & 1]) { item_multi[done] = 1; atomicAdd(n_multi, 1ULL); }
      acc_cnt[it & 1] = 0; acc_multi[it & 1] = 0;                   // used again two items from now, several barriers away
    }
  }
}

// Write pass of the items the match cache describes (every probe tuple has at most one match: unique build keys, or a semi-join): no
// table, no keys — each warp streams 32 cache words and probe row ids, ranks the hits with a ballot, claims its run of the item's output
// range with one shared-memory atomic and gathers the build row ids (the partition's 16 KB of row ids stay in L1 / L2 while its items run).
__global__ void __launch_bounds__(RJ_THREADS) k_rj_emit(const uint32_t* __restrict__ Rr, const uint32_t* __restrict__ Sr, const uint32_t* __restrict__ mcache,
                                                        const RjItem* __restrict__ items, const unsigned char* __restrict__ item_multi,
                                                        const unsigned long long* __restrict__ n_items_ptr, unsigned long long* tickets,
                                                        const unsigned long long* __restrict__ item_offsets, int32_t* __restrict__ outR, int32_t* __restrict__ outS,
                                                        const uint32_t* __restrict__ probe_payload, uint32_t probe_row_base, int carried_rows) {
  __shared__ TicketQueue2 tq;
  __shared__ uint32_t cursor[2];
  const long long n_items = (long long)*n_items_ptr;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const uint64_t pol = policy_evict_first();
  constexpr int U = 4;
  if (threadIdx.x < 2) cursor[threadIdx.x] = 0;
  long long item, next;
  ticket2_first(tickets, &tq, item, next);
  for (uint32_t it = 0; item < n_items; it++) {
    const long long pending = ticket_prefetch(tickets);
    if (next < n_items && !item_multi[next]) {                      // the next item's cache words and row ids into L2 while this one streams
      const RjItem wn = items[next];
      prefetch_l2_range(mcache + wn.s0, (wn.s1 - wn.s0) * 4u, RJ_THREADS);
      prefetch_l2_range(Sr + wn.s0, (wn.s1 - wn.s0) * 4u, RJ_THREADS);
    }
    if (!item_multi[item]) {                                        // uniform
      const RjItem w = items[item];
      const unsigned long long obase = item_offsets[item];
      uint32_t* cur = &cursor[it & 1];                              // double-buffered: the other one is reset for the next item below
      for (uint32_t j0 = w.s0 + (threadIdx.x & ~31u); j0 < w.s1; j0 += RJ_THREADS * U) {         // warp-uniform bounds
        uint32_t mc[U], srow[U], brow[U];
        #pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t j = j0 + u * RJ_THREADS + lane;
          mc[u] = j < w.s1 ? ld_stream<uint32_t>(mcache + j, pol) : ROW_NONE;
          srow[u] = j < w.s1 ? ld_stream<uint32_t>(Sr + j, pol) : 0u;
        }
        #pragma unroll
        for (int u = 0; u < U; u++) brow[u] = (outR && mc[u] != ROW_NONE) ? Rr[mc[u]] : 0u;
        #pragma unroll
        for (int u = 0; u < U; u++) {
          const bool hit = mc[u] != ROW_NONE;
          const unsigned hm = __ballot_sync(0xffffffffu, hit);
          if (hm) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cur, (uint32_t)__popc(hm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
              uint32_t prow = srow[u];
              if (!carried_rows) prow = probe_payload ? probe_payload[prow] : probe_row_base + prow;
              const unsigned long long pos = obase + base + __popc(hm & lt);
              if (outR) outR[pos] = (int32_t)brow[u];
              outS[pos] = (int32_t)prow;
            }
          }
        }
      }
    }
    if (threadIdx.x == 0) cursor[(it + 1) & 1] = 0;                  // nobody touches it during this item; ticket2_advance's barrier publishes the reset
    ticket2_advance(&tq, it, pending, item, next);
  }
}

template <typename Kern>
static unsigned rj_grid(Kern kern, size_t smem) {
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RJ_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (unsigned)(per_sm * sms);
}

template <typename K, bool WRITE>
static cudaError_t rj_launch(const RadixArea& r, const RadixScratch& s, unsigned long long* ticket, unsigned long long* item_totals, int32_t* outR, int32_t* outS,
                             const uint32_t* probe_payload, uint32_t probe_row_base, int carried_rows, int semi, uint32_t n_parts, uint32_t* mcache, unsigned long long* n_multi,
                             cudaStream_t stream) {
  auto kern = k_rj_join<K, WRITE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RjSmem<K>));
  if (e != cudaSuccess) return e;
  kern<<<rj_grid(kern, sizeof(RjSmem<K>)), RJ_THREADS, sizeof(RjSmem<K>), stream>>>((const K*)r.keys, r.rows, (const K*)s.a.keys, s.a.rows, s.items, s.item_start + n_parts, ticket,
                                                                                   item_totals, outR, outS, probe_payload, probe_row_base, carried_rows, semi, mcache, s.item_multi, n_multi);
  return cudaGetLastError();
}

// count: partition the probe relation like the table, make the items, count per item, scan. hdr_host: the table header as read back by
// the caller. item_totals: u64[radix_max_items(nS) + 1] (the scratch's offsets array); total_out: where the scan leaves the result size.
cudaError_t radix_count(const void* S, int64_t nS, int key_bytes, const TableHeader& hdr_host, const char* body, char* scratch_area,
                        unsigned long long* item_totals, unsigned long long* scan_block_sums, unsigned long long* ticket, unsigned long long* total_out,
                        uint32_t* mcache, unsigned long long* n_multi,
                        bool carry_rows, const uint32_t* probe_payload, uint32_t probe_row_base, bool semi, cudaStream_t stream) {
  const int b1 = (int)hdr_host.rj_bits1, b2 = (int)hdr_host.rj_bits2;
  const uint32_t n_parts = 1u << (b1 + b2);
  const RadixScratch s = radix_scratch(scratch_area, nS, key_bytes, b1, b2);
  RadixArea r;
  r.keys = const_cast<char*>(body) + hdr_host.rj_keys_off; r.rows = reinterpret_cast<uint32_t*>(const_cast<char*>(body) + hdr_host.rj_rows_off);
  r.offsets = reinterpret_cast<uint32_t*>(const_cast<char*>(body) + hdr_host.rj_offs_off);
  cudaError_t e = radix_partition2(S, carry_rows ? probe_payload : nullptr, carry_rows ? probe_row_base : 0u, nS, key_bytes, b1, b2, s.a.tmp_keys, s.a.tmp_rows, s.a.keys, s.a.rows,
                                   s.a.offsets, s.a.ws, s.a.ws_bytes, stream);
  if (e != cudaSuccess) return e;
  k_rj_item_counts<<<(n_parts + 255) / 256, 256, 0, stream>>>(r.offsets, s.a.offsets, n_parts, s.item_start);
  launch_scan(s.item_start, n_parts, s.item_start + n_parts + 1, nullptr, stream);
  k_rj_items<<<(n_parts + 255) / 256, 256, 0, stream>>>(r.offsets, s.a.offsets, n_parts, s.item_start, s.items);
  e = cudaMemsetAsync(item_totals, 0, (size_t)(s.max_items + 1) * 8, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(s.item_multi, 0, (size_t)s.max_items, stream);
  if (e != cudaSuccess) return e;
  e = key_bytes == 4 ? rj_launch<int32_t, false>(r, s, ticket, item_totals, nullptr, nullptr, nullptr, 0, 1, semi ? 1 : 0, n_parts, mcache, n_multi, stream)
                     : rj_launch<int64_t, false>(r, s, ticket, item_totals, nullptr, nullptr, nullptr, 0, 1, semi ? 1 : 0, n_parts, mcache, n_multi, stream);
  if (e != cudaSuccess) return e;
  launch_scan(item_totals, s.max_items, scan_block_sums, total_out, stream);
  return cudaGetLastError();
}

cudaError_t radix_write(int64_t nS, int key_bytes, const TableHeader& hdr_host, const char* body, char* scratch_area, unsigned long long* item_offsets,
                        unsigned long long* ticket, unsigned long long* ticket_emit, const uint32_t* mcache, unsigned long long* n_multi,
                        int32_t* outR, int32_t* outS, bool carried_rows, const uint32_t* probe_payload, uint32_t probe_row_base, bool semi, cudaStream_t stream) {
  const int b1 = (int)hdr_host.rj_bits1, b2 = (int)hdr_host.rj_bits2;
  const uint32_t n_parts = 1u << (b1 + b2);
  const RadixScratch s = radix_scratch(scratch_area, nS, key_bytes, b1, b2);
  RadixArea r;
  r.keys = const_cast<char*>(body) + hdr_host.rj_keys_off; r.rows = reinterpret_cast<uint32_t*>(const_cast<char*>(body) + hdr_host.rj_rows_off);
  r.offsets = reinterpret_cast<uint32_t*>(const_cast<char*>(body) + hdr_host.rj_offs_off);
  k_rj_emit<<<148 * 4, RJ_THREADS, 0, stream>>>(r.rows, s.a.rows, mcache, s.items, s.item_multi, s.item_start + n_parts, ticket_emit, item_offsets, outR, outS,
                                               probe_payload, probe_row_base, carried_rows ? 1 : 0);
  // the items the cache cannot describe (a probe tuple with several matches, partitions joined in rounds): the join again; exits at once when there are none
  return key_bytes == 4 ? rj_launch<int32_t, true>(r, s, ticket, item_offsets, outR, outS, probe_payload, probe_row_base, carried_rows ? 1 : 0, semi ? 1 : 0, n_parts, const_cast<uint32_t*>(mcache), n_multi, stream)
                        : rj_launch<int64_t, true>(r, s, ticket, item_offsets, outR, outS, probe_payload, probe_row_base, carried_rows ? 1 : 0, semi ? 1 : 0, n_parts, const_cast<uint32_t*>(mcache), n_multi, stream);
}

}  // namespace hj
