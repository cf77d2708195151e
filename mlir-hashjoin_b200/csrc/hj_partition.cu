// hj_partition.cu — K5: radix partition of (key, row id) tuples on a key hash, hand-written for sm_100a.
//
// New in the reference's terms: projectDescription.md:24 lists "partitioned hash-join" as left out. Three users here:
//   * the radix join of tables beyond L2 reach (hj_radix.cu): both relations are partitioned in TWO passes (<= 256 x 256 parts) until a
//     build partition fits a shared-memory table;
//   * hjPartition: one pass on an independent hash (which rank owns a key), local output, for the NCCL all-to-all plan;
//   * hjPartitionCount + hjPartitionPush: the same pass with the exchange fused in — every run is stored straight into the
//     receive buffer of the rank that owns it, through peer-mapped pointers (NVLink).
//
// One pass = three launches over BLOCKS of 8 192 .. 131 072 tuples (a block never straddles two segments of the previous pass):
//   k_rp_hist     per block: digit counts in shared memory -> mat[block][digit]; totals[segment][digit] by one atomic per (block, digit)
//   k_rp_scan     per (segment, 32 digits): first destination of every (block, digit) = segment base + exclusive scan of the digit
//                 totals + prefix over the segment's blocks   (coalesced 128-byte rows, two looks at the matrix)
//   k_rp_scatter  per block: tiles of 4 096 tuples (512 threads, 2 CTAs / SM) or 8 192 tuples (1 024 threads, more than 128 digits); a
//                 tile is ranked with ONE shared-memory atomic per tuple (ATOMS.ADD returns the rank inside the digit), staged
//                 digit-sorted in shared memory and written out run by run, so HBM / NVLink see contiguous stores (32 tuples per
//                 run at 256 digits). Three barriers per tile; the next tile's loads are issued before the output loop. No global
//                 atomics on the data path, no host round trip.
// The multi-GPU plans use hjPartitionPush twice over: with the peers' receive buffers as destinations (the exchange fused into this
// kernel), or with local staging buffers as destinations, the parts then crossing NVLink on the copy engines (dist.py).
// Indices are 32-bit (relations hold < 2^32 rows, join_v1.mlir:604-605: row ids are i32).
#include <algorithm>
#include <cstdio>
#include <type_traits>
#include "hj_common.cuh"
#include "hj_kernels.cuh"

namespace hj {

constexpr int RP_THREADS = 512;
constexpr int RP_ITEMS = 8;
constexpr int RP_TILE = RP_THREADS * RP_ITEMS;            // 4 096 tuples
// tuples per block: 2 .. 32 tiles, chosen so that a pass has about 2 000 blocks (148 SMs x 2 CTAs x ~7 waves): a block pays one cursor
// load and one un-overlapped first tile, so bigger is better until the tail of the grid shows
static inline uint32_t rp_block_tuples(int64_t n) {
  const int64_t tiles = (n + RP_TILE - 1) / RP_TILE;
  return (uint32_t)(std::min<int64_t>(32, std::max<int64_t>(2, tiles / 2000)) * RP_TILE);
}
constexpr int RP_MAX_FAN = 256;
constexpr int SEL_OWNER = 0, SEL_RADIX = 1, SEL_TABLE = 2, SEL_GROUP = 3;

struct RpBlock { uint32_t begin, end, seg, pad; };
struct DigitArgs { uint32_t fan, shift, mask; };

template <typename K, int SEL>
__device__ __forceinline__ uint32_t rp_digit(K key, const DigitArgs& da) {
  if (SEL == SEL_OWNER) return (uint32_t)(((uint64_t)KeyTraits<K>::part_hash(key) * da.fan) >> 32);      // any fan: which rank owns the key
  if (SEL == SEL_TABLE) return KeyTraits<K>::home_hash32(key) >> da.shift;                               // top bits of the hash that picks the bucket pair: a table slice
  if (SEL == SEL_GROUP) return KeyTraits<int64_t>::home_hash32((int64_t)key) >> da.shift;                // the grouped layout hashes the sign-extended key
  return (radix_hash<K>(key) >> da.shift) & da.mask;                                                    // a bit field of the radix hash
}

// ---------------------------------------------------------------------------------------------------------
// block descriptors: segment s = [seg_off[s], seg_off[s+1]) is cut into blocks of RP_BLOCK tuples (rp_block_tuples). One CTA.
// seg_off == nullptr: a single segment [0, n).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rp_blocks(const uint32_t* __restrict__ seg_off, uint32_t nseg, uint32_t n, uint32_t RP_BLOCK, RpBlock* __restrict__ blocks,
                                                   uint32_t* __restrict__ blk_start, uint32_t* __restrict__ n_blocks) {
  __shared__ uint32_t sm[33];
  __shared__ uint32_t first_sm[RP_MAX_FAN + 1], lo_sm[RP_MAX_FAN], hi_sm[RP_MAX_FAN];
  const uint32_t s = threadIdx.x;
  uint32_t lo = 0, hi = 0;
  if (s < nseg) { lo = seg_off ? seg_off[s] : 0u; hi = seg_off ? seg_off[s + 1] : n; }
  const uint32_t nb = (hi - lo + RP_BLOCK - 1) / RP_BLOCK;
  uint32_t total;
  const uint32_t first = block_exclusive_scan(nb, sm, &total);
  if (s < nseg) { first_sm[s] = first; lo_sm[s] = lo; hi_sm[s] = hi; blk_start[s] = first; }
  if (s == 0) { first_sm[nseg] = total; blk_start[nseg] = total; *n_blocks = total; }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < total; b += blockDim.x) {
    uint32_t a = 0, z = nseg;                                         // last segment whose first block is <= b (empty segments share a first block: take the last)
    while (z - a > 1) { const uint32_t m = (a + z) >> 1; if (first_sm[m] <= b) a = m; else z = m; }
    const uint32_t j = b - first_sm[a];
    const uint32_t begin = lo_sm[a] + j * RP_BLOCK;
    const uint32_t end = hi_sm[a] - begin < RP_BLOCK ? hi_sm[a] : begin + RP_BLOCK;
    blocks[b] = RpBlock{begin, end, a, 0u};
  }
}

// ---------------------------------------------------------------------------------------------------------
// pass 1 of 2: histogram
// ---------------------------------------------------------------------------------------------------------
template <typename K, int SEL>
__global__ void __launch_bounds__(RP_THREADS) k_rp_hist(const K* __restrict__ keys, const RpBlock* __restrict__ blocks, const uint32_t* __restrict__ n_blocks,
                                                        DigitArgs da, uint32_t* __restrict__ mat, uint32_t* __restrict__ totals) {
  __shared__ uint32_t cnt[RP_MAX_FAN];
  const uint32_t b = blockIdx.x;
  if (b >= *n_blocks) return;
  const RpBlock blk = blocks[b];
  if (threadIdx.x < RP_MAX_FAN) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t pol = policy_evict_first();
  for (uint32_t i0 = blk.begin; i0 < blk.end; i0 += RP_TILE) {
    K key[RP_ITEMS];
    #pragma unroll
    for (int e = 0; e < RP_ITEMS; e++) { const uint32_t i = i0 + e * RP_THREADS + threadIdx.x; key[e] = i < blk.end ? ld_stream<K>(keys + i, pol) : K(0); }
    #pragma unroll
    for (int e = 0; e < RP_ITEMS; e++) if (i0 + e * RP_THREADS + threadIdx.x < blk.end) atomicAdd(&cnt[rp_digit<K, SEL>(key[e], da)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < da.fan) {
    const uint32_t c = cnt[threadIdx.x];
    mat[(size_t)b * da.fan + threadIdx.x] = c;
    if (c) atomicAdd(&totals[(size_t)blk.seg * da.fan + threadIdx.x], c);
  }
}

// ---------------------------------------------------------------------------------------------------------
// scan: mat[block][digit] := first destination element of that (block, digit)
//   local:  seg_base[s] + exclusive scan of totals[s][.] + prefix over the blocks of s;  offsets[s * fan + d] = start of part (s, d)
//   push:   start_in[d] (this rank's cursor in the owner's receive buffer) + prefix over the blocks
// grid = (ceil(fan / 32), nseg), 1024 threads: lane = digit, warp = a contiguous run of the segment's blocks.
// ---------------------------------------------------------------------------------------------------------
constexpr int RPSCAN_THREADS = 1024;
__global__ void __launch_bounds__(RPSCAN_THREADS) k_rp_scan(uint32_t* __restrict__ mat, const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ seg_off, uint32_t n,
                                                            uint32_t fan, const uint32_t* __restrict__ totals, uint32_t* __restrict__ offsets,
                                                            const unsigned long long* __restrict__ start_in, const uint32_t* __restrict__ n_blocks) {
  __shared__ uint32_t sm[33];
  __shared__ uint32_t ex[RP_MAX_FAN];
  __shared__ uint32_t wsum[32][33];
  const uint32_t s = blockIdx.y, nseg = gridDim.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t seg_base = seg_off ? seg_off[s] : 0u;
  {
    uint32_t all;
    const uint32_t e = block_exclusive_scan(threadIdx.x < fan ? totals[(size_t)s * fan + threadIdx.x] : 0u, sm, &all);
    if (threadIdx.x < fan) {
      ex[threadIdx.x] = e;
      if (offsets && blockIdx.x == 0) {
        offsets[(size_t)s * fan + threadIdx.x] = seg_base + e;
        if (s == nseg - 1 && threadIdx.x == fan - 1) offsets[(size_t)nseg * fan] = seg_off ? seg_off[nseg] : n;
      }
    }
  }
  __syncthreads();
  const uint32_t d = blockIdx.x * 32 + lane;
  const uint32_t b0 = blk_start[s], b1 = blk_start[s + 1];
  const uint32_t nrows = b1 - b0, per_warp = (nrows + 31) / 32, r0 = warp * per_warp;
  const uint32_t rows = r0 < nrows ? (nrows - r0 < per_warp ? nrows - r0 : per_warp) : 0;
  uint32_t sum = 0;
  if (d < fan) {
    #pragma unroll 8
    for (uint32_t i = 0; i < rows; i++) sum += mat[(size_t)(b0 + r0 + i) * fan + d];
  }
  wsum[warp][lane] = sum;
  __syncthreads();
  if (d < fan) {
    uint32_t run = start_in ? (uint32_t)start_in[d] : seg_base + ex[d];
    for (int w = 0; w < warp; w++) run += wsum[w][lane];
    for (uint32_t i0 = 0; i0 < rows; i0 += 8) {
      uint32_t v[8];
      #pragma unroll
      for (int i = 0; i < 8; i++) v[i] = i0 + i < rows ? mat[(size_t)(b0 + r0 + i0 + i) * fan + d] : 0u;
      #pragma unroll
      for (int i = 0; i < 8; i++) if (i0 + i < rows) { mat[(size_t)(b0 + r0 + i0 + i) * fan + d] = run; run += v[i]; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// pass 2 of 2: scatter
// ---------------------------------------------------------------------------------------------------------
template <typename K, bool PUSH, int THREADS, int ITEMS>
struct RpSmem {
  static constexpr int TILE = THREADS * ITEMS;
  K skeys[TILE];
  uint32_t srows[TILE];
  uint32_t cnt[RP_MAX_FAN];                   // tuples of each digit in the tile (rank counter), zero between tiles
  uint32_t lbase[RP_MAX_FAN];                 // first staged position of each digit
  uint32_t delta[RP_MAX_FAN];                 // destination index of staged position i of digit d: delta[d] + i (mod 2^32)
  uint32_t gcur[RP_MAX_FAN];                  // running destination cursor of each digit for this block
  K* kptr[PUSH ? RP_MAX_FAN : 1];             // per-digit destination buffers: only the push into peers' receive buffers has more than one
  uint32_t* rptr[PUSH ? RP_MAX_FAN : 1];
  unsigned char sdig[PUSH ? TILE : 1];        // push only: the owner of every staged tuple (the local variants hash again in the output loop)
};

template <typename K, int SEL, bool PUSH, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, THREADS >= 1024 ? 1 : 1024 / THREADS) k_rp_scatter(const K* __restrict__ keys, const uint32_t* __restrict__ rows, uint32_t row_base,
                                                              const RpBlock* __restrict__ blocks, const uint32_t* __restrict__ n_blocks, DigitArgs da,
                                                              K* __restrict__ out_keys, uint32_t* __restrict__ out_rows,
                                                              K* const* __restrict__ dst_keys, uint32_t* const* __restrict__ dst_rows,
                                                              const uint32_t* __restrict__ mat) {
  extern __shared__ __align__(16) unsigned char rp_raw[];
  RpSmem<K, PUSH, THREADS, ITEMS>& sm = *reinterpret_cast<RpSmem<K, PUSH, THREADS, ITEMS>*>(rp_raw);
  constexpr int RP_THREADS = THREADS, RP_ITEMS = ITEMS, RP_TILE = THREADS * ITEMS;   // shadow the file-scope geometry (that of k_rp_hist)
  const uint32_t b = blockIdx.x;
  if (b >= *n_blocks) return;
  const RpBlock blk = blocks[b];
  const uint32_t fan = da.fan;
  for (uint32_t d = threadIdx.x; d < RP_MAX_FAN; d += THREADS) {
    sm.cnt[d] = 0;
    sm.gcur[d] = d < fan ? mat[(size_t)b * fan + d] : 0u;
    if (PUSH && d < fan) { sm.kptr[d] = dst_keys[d]; sm.rptr[d] = dst_rows[d]; }
  }
  const uint64_t pol = policy_evict_first();
  K key[RP_ITEMS]; uint32_t row[RP_ITEMS];
  // Tile t+1 is loaded into registers right before tile t's output loop and tile t+2 is pulled into L2 at the same time, so the register
  // loads are L2 hits. Full tiles (all but the last of a block) run a copy of the code without per-element bounds tests.
  // (Measured and rejected: keys only in registers, row ids re-read in the stage phase, 40 registers for a third CTA per SM — 2^28 i64
  // tuples per pass 2.13 -> 3.23 ms: the spills and the exposed row loads cost more than the occupancy gains.)
  auto load_tile = [&](uint32_t base, auto full_tag) {       // element e of this thread sits at base + e * RP_THREADS + tid: every load instruction is one
    constexpr bool FULL = decltype(full_tag)::value;         // contiguous run per warp whatever the alignment of `base` (segments start anywhere)
    #pragma unroll
    for (int e = 0; e < RP_ITEMS; e++) {
      const uint32_t i = base + e * RP_THREADS + threadIdx.x;
      key[e] = (FULL || i < blk.end) ? ld_stream<K>(keys + i, pol) : K(0);
      row[e] = rows ? ((FULL || i < blk.end) ? ld_stream<uint32_t>(rows + i, pol) : 0u) : row_base + i;
    }
    constexpr int KEYS_PER_SECTOR = 32 / sizeof(K), KEY_SECTORS = RP_TILE / KEYS_PER_SECTOR;   // the tile after: one prefetch per 32-byte sector
    #pragma unroll
    for (int q = 0; q < KEY_SECTORS / RP_THREADS; q++) {
      const uint32_t ahead = base + RP_TILE + (q * RP_THREADS + threadIdx.x) * KEYS_PER_SECTOR;
      if (ahead < blk.end) prefetch_l2(keys + ahead);
    }
    if (rows && base + RP_TILE + threadIdx.x * 8 < blk.end) prefetch_l2(rows + base + RP_TILE + threadIdx.x * 8);
  };
  auto load_next = [&](uint32_t base) {
    if (base + RP_TILE <= blk.end) load_tile(base, std::true_type{}); else load_tile(base, std::false_type{});
  };
  auto tile_body = [&](uint32_t base, uint32_t count, auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    uint32_t pr[RP_ITEMS];                                   // digit << 16 | rank inside the digit
    #pragma unroll
    for (int e = 0; e < RP_ITEMS; e++) {
      pr[e] = 0xFFFFFFFFu;
      if (FULL || e * RP_THREADS + threadIdx.x < count) {
        const uint32_t d = rp_digit<K, SEL>(key[e], da);
        pr[e] = (d << 16) | atomicAdd(&sm.cnt[d], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                  // exclusive scan of the <= 256 digit counts (8 per lane); cursors advance, counters reset
      uint32_t v[RP_MAX_FAN / 32], sum = 0;
      #pragma unroll
      for (int q = 0; q < RP_MAX_FAN / 32; q++) { v[q] = sm.cnt[threadIdx.x * (RP_MAX_FAN / 32) + q]; sum += v[q]; }
      uint32_t run = warp_inclusive_scan(sum) - sum;
      #pragma unroll
      for (int q = 0; q < RP_MAX_FAN / 32; q++) {
        const int d = threadIdx.x * (RP_MAX_FAN / 32) + q;
        sm.lbase[d] = run; const uint32_t g = sm.gcur[d]; sm.delta[d] = g - run; sm.gcur[d] = g + v[q]; sm.cnt[d] = 0;
        run += v[q];
      }
    }
    __syncthreads();
    #pragma unroll
    for (int e = 0; e < RP_ITEMS; e++) {
      if (FULL || pr[e] != 0xFFFFFFFFu) {
        const uint32_t d = pr[e] >> 16, pos = sm.lbase[d] + (pr[e] & 0xFFFFu);
        sm.skeys[pos] = key[e]; sm.srows[pos] = row[e];
        if (PUSH) sm.sdig[pos] = (unsigned char)d;
      }
    }
    __syncthreads();
    if (base + RP_TILE < blk.end) load_next(base + RP_TILE);  // in flight while this tile's runs are written
    #pragma unroll 2
    for (int it = 0; it < RP_ITEMS; it++) {                   // digit-sorted: consecutive i -> consecutive destination addresses
      const uint32_t i = it * RP_THREADS + threadIdx.x;
      if (FULL || i < count) {
        const K k = sm.skeys[i];
        const uint32_t r = sm.srows[i], d = PUSH ? (uint32_t)sm.sdig[i] : rp_digit<K, SEL>(k, da);   // local: hashed again rather than staged (the shared-memory pipe is the limiter)
        const uint32_t idx = sm.delta[d] + i;                            // (ncu: mio_throttle + short_scoreboard 15 warps per issue slot at 25 % issue), ALU is free
        if (PUSH) { sm.kptr[d][idx] = k; sm.rptr[d][idx] = r; }
        else { out_keys[idx] = k; out_rows[idx] = r; }
      }
    }
    // the next tile's rank phase only touches cnt (reset above); its scan and stage phases come after barriers every thread reaches
    // only once it has left this output loop
  };
  load_next(blk.begin);
  __syncthreads();
  for (uint32_t base = blk.begin; base < blk.end; base += RP_TILE) {
    if (blk.end - base >= (uint32_t)RP_TILE) tile_body(base, (uint32_t)RP_TILE, std::true_type{});
    else tile_body(base, blk.end - base, std::false_type{});
  }
}

// (Measured and removed — git history has the kernel: a warp-specialised variant, 512 producer threads that load / rank / stage 8 192-tuple
// tiles into two shared-memory buffers and 512 consumer threads that write the other buffer's runs, named-barrier hand-over, setmaxnreg
// 96 / 32. Overlapping the phases buys nothing because the shared-memory pipe is busy in ALL of them — rank atomics, conflicting stage
// stores, output loads (tools/membench5.cu: rank + stage alone are 5 400 of a tile's 14 600 cycles): 2^28 x 2^28 i64 join 11.8 -> 12.8 ms,
// configs 2-sparse and 4 unchanged. Ballot / match_any ranking instead of the returning atomic: 2.6x / 4x slower, same microbenchmark.)
// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static inline int64_t r256(int64_t x) { return (x + 255) / 256 * 256; }
static inline int64_t rp_max_blocks(int64_t n, int nseg) { const int64_t b = rp_block_tuples(n); return (n + b - 1) / b + nseg + 1; }

// workspace of one pass: [n_blocks u32 (256 B)] [blk_start u32 x (nseg + 1)] [totals u32 x nseg x fan] [blocks] [mat u32 x blocks x fan]
struct RpWorkspace { uint32_t* n_blocks; uint32_t* blk_start; uint32_t* totals; RpBlock* blocks; uint32_t* mat; int64_t max_blocks; };
int64_t rp_workspace_bytes(int64_t n, int nseg, int fan) {
  const int64_t nb = rp_max_blocks(n, nseg);
  return 256 + r256((int64_t)(nseg + 1) * 4) + r256((int64_t)nseg * fan * 4) + r256(nb * (int64_t)sizeof(RpBlock)) + r256(nb * fan * 4);
}
static RpWorkspace rp_workspace(void* ws, int64_t n, int nseg, int fan) {
  RpWorkspace w;
  char* p = reinterpret_cast<char*>(ws);
  w.max_blocks = rp_max_blocks(n, nseg);
  w.n_blocks = reinterpret_cast<uint32_t*>(p); p += 256;
  w.blk_start = reinterpret_cast<uint32_t*>(p); p += r256((int64_t)(nseg + 1) * 4);
  w.totals = reinterpret_cast<uint32_t*>(p); p += r256((int64_t)nseg * fan * 4);
  w.blocks = reinterpret_cast<RpBlock*>(p); p += r256(w.max_blocks * (int64_t)sizeof(RpBlock));
  w.mat = reinterpret_cast<uint32_t*>(p);
  return w;
}

template <typename K, int SEL, bool PUSH, int THREADS, int ITEMS = 8>
static cudaError_t rp_launch_scatter_t(const void* keys, const uint32_t* rows, uint32_t row_base, const RpWorkspace& w, DigitArgs da, void* out_keys, uint32_t* out_rows,
                                       void* const* dst_keys, uint32_t* const* dst_rows, cudaStream_t stream) {
  auto kern = k_rp_scatter<K, SEL, PUSH, THREADS, ITEMS>;
  using Smem = RpSmem<K, PUSH, THREADS, ITEMS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));     // cheap, per device: no cached flag
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)w.max_blocks, THREADS, sizeof(Smem), stream>>>((const K*)keys, rows, row_base, w.blocks, w.n_blocks, da, (K*)out_keys, out_rows,
                                                                (K* const*)dst_keys, dst_rows, w.mat);
  return cudaGetLastError();
}
// CTA shape of the scatter kernel. What decides is the RUN LENGTH (tile / fan): a run is written as one contiguous piece at an arbitrary
// alignment, and the shorter it is, the more partial 32-byte sectors reach L2 and DRAM. Measured on 2^28 i64 tuples, fan 256 (ncu, per pass):
//   256 threads, 2 048-tuple tiles,  8-tuple runs, 4 CTAs / SM: 3.33 ms, DRAM 3.6 GB read + 4.5 GB written (5.3 GB algorithmic)
//   512 threads, 4 096-tuple tiles, 16-tuple runs, 2 CTAs / SM: 2.13 ms, DRAM 2.8 + 3.7 GB
//  1024 threads, 8 192-tuple tiles, 32-tuple runs, 1 CTA  / SM: ~1.8 ms (bench: 2^28 x 2^28 i64 join 14.6 -> 12.6 ms)
// With fan <= 128 the 512-thread shape already has 32-tuple runs and its two CTAs per SM overlap each other's barriers (config 2 with
// sparse keys, fan 64: 5.09 ms vs 5.22 ms with 1024 threads). 16 tuples per thread instead of 8 (longer runs at the same CTA shape) spills:
// 112 bytes of stack per thread for i32 keys, 184 for i64, at the 64 registers two 512-thread CTAs leave. 0 = choose by fan.
static int g_rp_threads = 0;
void set_partition_threads(int t) { g_rp_threads = t; }
template <typename K, int SEL, bool PUSH>
static cudaError_t rp_launch_scatter(const void* keys, const uint32_t* rows, uint32_t row_base, const RpWorkspace& w, DigitArgs da, void* out_keys, uint32_t* out_rows,
                                     void* const* dst_keys, uint32_t* const* dst_rows, cudaStream_t stream) {
  const int threads = g_rp_threads ? g_rp_threads : (da.fan > 128 ? 1024 : 512);
  if (threads == 1024) return rp_launch_scatter_t<K, SEL, PUSH, 1024>(keys, rows, row_base, w, da, out_keys, out_rows, dst_keys, dst_rows, stream);
  if (threads == 256) return rp_launch_scatter_t<K, SEL, PUSH, 256>(keys, rows, row_base, w, da, out_keys, out_rows, dst_keys, dst_rows, stream);
  return rp_launch_scatter_t<K, SEL, PUSH, 512>(keys, rows, row_base, w, da, out_keys, out_rows, dst_keys, dst_rows, stream);
}

template <typename K, int SEL>
static cudaError_t rp_hist(const void* keys, int64_t n, const uint32_t* seg_off, int nseg, DigitArgs da, const RpWorkspace& w, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(w.totals, 0, (size_t)nseg * da.fan * 4, stream);
  if (e != cudaSuccess) return e;
  k_rp_blocks<<<1, 256, 0, stream>>>(seg_off, (uint32_t)nseg, (uint32_t)n, rp_block_tuples(n), w.blocks, w.blk_start, w.n_blocks);
  k_rp_hist<K, SEL><<<(unsigned)w.max_blocks, RP_THREADS, 0, stream>>>((const K*)keys, w.blocks, w.n_blocks, da, w.mat, w.totals);
  return cudaGetLastError();
}

// One local pass: segments of the input (seg_off: nseg + 1 device offsets, nullptr = one segment) are each split into `fan` parts by the
// digit `da` selects; offsets[s * fan + d] = first output element of part (s, d), offsets[nseg * fan] = n.
template <typename K, int SEL>
static cudaError_t rp_pass(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, const uint32_t* seg_off, int nseg, DigitArgs da,
                           void* out_keys, uint32_t* out_rows, uint32_t* offsets, void* ws, cudaStream_t stream) {
  const RpWorkspace w = rp_workspace(ws, n, nseg, (int)da.fan);
  cudaError_t e = rp_hist<K, SEL>(keys, n, seg_off, nseg, da, w, stream);
  if (e != cudaSuccess) return e;
  k_rp_scan<<<dim3((da.fan + 31) / 32, (unsigned)nseg), RPSCAN_THREADS, 0, stream>>>(w.mat, w.blk_start, seg_off, (uint32_t)n, da.fan, w.totals, offsets, nullptr, w.n_blocks);
  return rp_launch_scatter<K, SEL, false>(keys, rows, row_base, w, da, out_keys, out_rows, nullptr, nullptr, stream);
}

// ---- two-level radix partition (hj_radix.cu) ----------------------------------------------------------------
int64_t radix_partition2_workspace_bytes(int64_t n, int bits1, int bits2) {
  return r256(((int64_t)(1 << bits1) + 1) * 4) + std::max(rp_workspace_bytes(n, 1, 1 << bits1), rp_workspace_bytes(n, 1 << bits1, 1 << bits2));
}
cudaError_t radix_partition2(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int bits1, int bits2,
                             void* tmp_keys, uint32_t* tmp_rows, void* out_keys, uint32_t* out_rows, uint32_t* offsets, void* ws, int64_t ws_bytes,
                             cudaStream_t stream) {
  if (bits1 < 1 || bits1 > 8 || bits2 < 0 || bits2 > 8 || n < 0 || n > 0xFFFFFFFFLL || ws_bytes < radix_partition2_workspace_bytes(n, bits1, bits2)) return cudaErrorInvalidValue;
  const int P1 = 1 << bits1, P2 = 1 << bits2;
  uint32_t* off1 = reinterpret_cast<uint32_t*>(ws);                                     // level-1 part offsets (P1 + 1), kept for pass 2
  void* ws_pass = reinterpret_cast<char*>(ws) + r256(((int64_t)P1 + 1) * 4);
  const DigitArgs d1{(uint32_t)P1, (uint32_t)(32 - bits1), (uint32_t)(P1 - 1)};
  const DigitArgs d2{(uint32_t)P2, (uint32_t)(32 - bits1 - bits2), (uint32_t)(P2 - 1)};
  cudaError_t e;
  if (bits2 == 0) {
    if (key_bytes == 4) return rp_pass<int32_t, SEL_RADIX>(keys, rows, row_base, n, nullptr, 1, d1, out_keys, out_rows, offsets, ws_pass, stream);
    return rp_pass<int64_t, SEL_RADIX>(keys, rows, row_base, n, nullptr, 1, d1, out_keys, out_rows, offsets, ws_pass, stream);
  }
  if (key_bytes == 4) {
    e = rp_pass<int32_t, SEL_RADIX>(keys, rows, row_base, n, nullptr, 1, d1, tmp_keys, tmp_rows, off1, ws_pass, stream);
    if (e == cudaSuccess) e = rp_pass<int32_t, SEL_RADIX>(tmp_keys, tmp_rows, 0, n, off1, P1, d2, out_keys, out_rows, offsets, ws_pass, stream);
  } else {
    e = rp_pass<int64_t, SEL_RADIX>(keys, rows, row_base, n, nullptr, 1, d1, tmp_keys, tmp_rows, off1, ws_pass, stream);
    if (e == cudaSuccess) e = rp_pass<int64_t, SEL_RADIX>(tmp_keys, tmp_rows, 0, n, off1, P1, d2, out_keys, out_rows, offsets, ws_pass, stream);
  }
  return e;
}

// ---- one pass on the TABLE hash: slice s of the output holds the tuples whose bucket pairs lie in the s-th 2^-bits of the inline table ----
int64_t slice_partition_workspace_bytes(int64_t n, int bits) { return rp_workspace_bytes(n, 1, 1 << bits); }
cudaError_t slice_partition(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int bits, bool grouped,
                            void* out_keys, uint32_t* out_rows, uint32_t* offsets, void* ws, int64_t ws_bytes, cudaStream_t stream) {
  if (bits < 1 || bits > 8 || n < 0 || n > 0xFFFFFFFFLL || ws_bytes < slice_partition_workspace_bytes(n, bits)) return cudaErrorInvalidValue;
  const DigitArgs da{(uint32_t)1 << bits, (uint32_t)(32 - bits), (uint32_t)((1 << bits) - 1)};
  if (key_bytes == 4) return grouped ? rp_pass<int32_t, SEL_GROUP>(keys, rows, row_base, n, nullptr, 1, da, out_keys, out_rows, offsets, ws, stream)
                                     : rp_pass<int32_t, SEL_TABLE>(keys, rows, row_base, n, nullptr, 1, da, out_keys, out_rows, offsets, ws, stream);
  return rp_pass<int64_t, SEL_TABLE>(keys, rows, row_base, n, nullptr, 1, da, out_keys, out_rows, offsets, ws, stream);      // i64: both layouts hash alike
}

// ---- one pass on the owner hash: hjPartition / hjPartitionCount / hjPartitionPush -----------------------------
// workspace: [pass workspace] [offsets u32 x (n_parts + 1)] [counts u64 x n_parts]
int64_t partition_workspace_bytes(int64_t n, int n_parts) { return rp_workspace_bytes(n, 1, n_parts) + r256(((int64_t)n_parts + 1) * 4) + r256((int64_t)n_parts * 8); }

__global__ void k_rp_widen(const uint32_t* __restrict__ in, unsigned long long* __restrict__ out, int count) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) out[i] = in[i];
}

cudaError_t radix_partition(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int n_parts,
                            void* out_keys, uint32_t* out_rows, unsigned long long* offsets, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (n_parts < 1 || n_parts > RP_MAX_FAN || n < 0 || n > 0xFFFFFFFFLL || workspace_bytes < partition_workspace_bytes(n, n_parts)) return cudaErrorInvalidValue;
  uint32_t* off32 = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + rp_workspace_bytes(n, 1, n_parts));
  const DigitArgs da{(uint32_t)n_parts, 0u, 0u};
  cudaError_t e = key_bytes == 4 ? rp_pass<int32_t, SEL_OWNER>(keys, rows, row_base, n, nullptr, 1, da, out_keys, out_rows, off32, workspace, stream)
                                 : rp_pass<int64_t, SEL_OWNER>(keys, rows, row_base, n, nullptr, 1, da, out_keys, out_rows, off32, workspace, stream);
  if (e != cudaSuccess) return e;
  k_rp_widen<<<1, 256, 0, stream>>>(off32, offsets, n_parts + 1);
  return cudaGetLastError();
}

// histogram only: counts[p] (device, u64) = tuples of partition p; the per-block matrix stays in the workspace for partition_push
cudaError_t partition_count(const void* keys, int64_t n, int key_bytes, int n_parts, unsigned long long* counts, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (n_parts < 1 || n_parts > RP_MAX_FAN || n < 0 || n > 0xFFFFFFFFLL || workspace_bytes < partition_workspace_bytes(n, n_parts)) return cudaErrorInvalidValue;
  const RpWorkspace w = rp_workspace(workspace, n, 1, n_parts);
  const DigitArgs da{(uint32_t)n_parts, 0u, 0u};
  cudaError_t e = key_bytes == 4 ? rp_hist<int32_t, SEL_OWNER>(keys, n, nullptr, 1, da, w, stream) : rp_hist<int64_t, SEL_OWNER>(keys, n, nullptr, 1, da, w, stream);
  if (e != cudaSuccess) return e;
  if (counts) k_rp_widen<<<1, 256, 0, stream>>>(w.totals, counts, n_parts);
  return cudaGetLastError();
}

// Fused partition + exchange: peer_keys[p] / peer_rows[p] are DEVICE arrays of peer-mapped receive-buffer pointers, cursors[p] holds the
// first element of this rank's region in partition p's receive buffer (from the all-gathered count matrix). Must follow
// partition_count() on the same keys and workspace.
cudaError_t partition_push(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int n_parts,
                           void* const* peer_keys, uint32_t* const* peer_rows, const unsigned long long* cursors, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream) {
  if (n_parts < 1 || n_parts > RP_MAX_FAN || n < 0 || n > 0xFFFFFFFFLL || workspace_bytes < partition_workspace_bytes(n, n_parts)) return cudaErrorInvalidValue;
  const RpWorkspace w = rp_workspace(workspace, n, 1, n_parts);
  const DigitArgs da{(uint32_t)n_parts, 0u, 0u};
  k_rp_scan<<<dim3((n_parts + 31) / 32, 1), RPSCAN_THREADS, 0, stream>>>(w.mat, w.blk_start, nullptr, (uint32_t)n, (uint32_t)n_parts, w.totals, nullptr, cursors, w.n_blocks);
  if (key_bytes == 4) return rp_launch_scatter<int32_t, SEL_OWNER, true>(keys, rows, row_base, w, da, nullptr, nullptr, peer_keys, peer_rows, stream);
  return rp_launch_scatter<int64_t, SEL_OWNER, true>(keys, rows, row_base, w, da, nullptr, nullptr, peer_keys, peer_rows, stream);
}

}  // namespace hj
