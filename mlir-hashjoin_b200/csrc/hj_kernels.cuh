// hj_kernels.cuh — host-side launch interface of the sm_100a join kernels (implemented in hj_kernels.cu).
// Everything here takes DEVICE pointers and an explicit stream; nothing synchronises except count_rows().
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hj {

constexpr int HEADER_BYTES = 256;              // table workspace = header + body
// Most parts one radix-partition pass makes. 256 parts leave 33 MB table slices for 2^28 i64 build rows, with two slices live at a
// time (50 % L2 hits in k_build_hash). 384 parts (22 MB slices) measured WORSE: 2^28 x 2^28 i64 20.6 ms vs 19.6 ms — shorter runs
// in the scatter kernel cost more than the slices' better residency gains.
constexpr int PART_MAX = 256;

// tile geometry shared by count/scan/write (must agree between the two probe passes)
constexpr int BLOCK_THREADS = 256;
constexpr int VECS_PER_THREAD = 2;
// consecutive tiles one CTA owns (the scan runs over chunks). i64 keys: one tile per CTA — a big table is probed in slice
// order, and the rows in flight (CTAs x chunk) times 32 table bytes per row must stay within L2 reach.
__host__ __device__ constexpr int chunk_tiles(int key_bytes) { return key_bytes == 4 ? 8 : 1; }
__host__ __device__ constexpr int keys_per_vec(int key_bytes) { return 16 / key_bytes; }
__host__ __device__ constexpr int tile_keys(int key_bytes) { return BLOCK_THREADS * VECS_PER_THREAD * keys_per_vec(key_bytes); }
__host__ __device__ constexpr int chunk_keys(int key_bytes) { return chunk_tiles(key_bytes) * tile_keys(key_bytes); }

int64_t preferred_pairs(int64_t n_rows, int key_bytes);
int64_t table_bytes(int64_t n_rows, int key_bytes);
int64_t scratch_bytes(int64_t n_probe, int key_bytes);
int64_t num_chunks(int64_t n_probe, int key_bytes);

// Scratch layout (device): [ match cache: u32 x round_up(n_probe, chunk) ][ as much again ][ chunk offsets: u64 x (nchunks + 1) ]
// Selective joins (few probe rows hit) use the first two areas together as per-chunk HIT LISTS instead: chunk c's hits are the
// first cnt[c] 8-byte entries (matched build row, position of the probe row inside the chunk) of its slice, see k_count_sparse.
struct ScratchView {
  uint32_t* mcache;
  uint2* hit_list;
  uint32_t* run_start;                // grouped layout: first row-id slot of each probe row's run (the second half of the cache area)
  uint32_t* warp_counts;              // hit-list mode: entries in each of the 8 warp lists of a chunk
  unsigned long long* chunk_offsets;  // after scan: exclusive offsets; [nchunks] = total
  int64_t nchunks;
  unsigned long long* counters;       // [0..2] ticket counters of the bounded-grid probe kernels, [3] hit-list mode flag, [4] ticket of k_count_sparse
  char* reorder;                      // slice-ordered copy of the probe relation (big tables only)
};
ScratchView scratch_view(void* scratch, int64_t n_probe, int key_bytes);

void set_allow_dense(int on);
void set_locality(int on);
void set_sparse(int policy);    // hit lists for selective joins: 0 never, 1 decided on the device from a sample of the probe keys (default), 2 always
void set_dup_sample(int on);    // 1 (default): sample the build keys for duplicates before attempting the inline layout
void set_dense_waves(int k);    // grid of the direct-address probe kernels: 0 = one chunk per CTA (default), k = at most k resident waves
void set_tma_count(int on);     // debug/bench switch: 0 = LDG/STG streams in the direct-address count kernel instead of TMA bulk copies      // debug/bench switch: 0 disables the slice-ordered build/probe of big tables
bool table_is_big(int64_t n_rows, int key_bytes);   // debug/bench switch: 0 forces the hash layout even for dense key ranges

// K0+K1: clear + build.  payload == nullptr -> row id = row_base + i  (join_v1.mlir:232 stores the thread index).
cudaError_t build_table(const void* R, int64_t nR, int key_bytes, const uint32_t* payload, uint32_t row_base,
                        void* table, int64_t table_bytes_, cudaStream_t stream);
// K2+K3: count + scan (async).  Total lands in chunk_offsets[nchunks].  big_hint: the table may be beyond L2 reach (then the
// header is read back once and, unless the layout is direct-address, the probe relation is reordered by table slice first);
// *reordered tells write_pairs which copy of the relation the match cache refers to.
// carry_rows: the probe row ids (payload column or row base) are known now, so a slice-ordered copy carries THEM (REORDER_ROWS)
// instead of the original index (REORDER_INDEX); write_pairs must then be given the same ids or none.
constexpr int REORDER_NONE = 0, REORDER_INDEX = 1, REORDER_ROWS = 2;
// range_hint: the table may have been built under hjSetAllowDense(2) (count by range): queue k_count_range / k_write_range too.
int allow_dense();
cudaError_t count_rows_async(const void* S, int64_t nS, int key_bytes, const void* table, void* scratch, bool big_hint, bool range_hint, int* reordered,
                             bool carry_rows, const uint32_t* probe_payload, uint32_t probe_row_base, cudaStream_t stream);
// K4: write pairs.
cudaError_t write_pairs(const void* S, int64_t nS, int key_bytes, const void* table, const void* scratch,
                        int32_t* outR, int32_t* outS, const uint32_t* probe_payload, uint32_t probe_row_base, int reordered, bool range_hint,
                        cudaStream_t stream);

// K2+K3+K4 fused (single pass, decoupled look-back): unique layouts only; the total lands where count_rows_async puts it.
cudaError_t join_fused_async(const void* S, int64_t nS, int key_bytes, const void* table, void* scratch, int32_t* outR, int32_t* outS, int64_t capacity,
                             const uint32_t* probe_payload, uint32_t probe_row_base, cudaStream_t stream);
cudaError_t read_table_mode(const void* table, uint32_t* mode, uint32_t* all_present, cudaStream_t stream);

// K5: radix partition by the key hash (multi-GPU shuffle feed).  Two launches: histogram, scatter.
//   counts: u64[n_parts] (device, zeroed by the call); offsets computed on device; keys/rows scattered so that
//   partition p occupies [offsets[p], offsets[p+1]) of out_keys/out_rows.  offsets: u64[n_parts+1] device.
cudaError_t radix_partition(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int n_parts,
                            void* out_keys, uint32_t* out_rows, unsigned long long* offsets, void* workspace, int64_t workspace_bytes,
                            int sel, cudaStream_t stream);          // sel 0: owner hash (multi-GPU), 1/2: table-slice hash (inline/grouped)
int64_t partition_workspace_bytes(int64_t n, int n_parts);
cudaError_t partition_count(const void* keys, int64_t n, int key_bytes, int n_parts, unsigned long long* counts, void* workspace, int64_t workspace_bytes,
                            int sel, cudaStream_t stream);
cudaError_t partition_push(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int n_parts,
                           void* const* peer_keys, uint32_t* const* peer_rows, const unsigned long long* cursors, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream);

// K6: verification helpers — order-independent digest of a pair stream: out[0] += sum(mix64(pair)), out[1] ^= xor.
cudaError_t pair_digest(const int32_t* outR, const int32_t* outS, int64_t n, unsigned long long* out2, cudaStream_t stream);

// seeded generators, bit-identical to oracle/oracle_join.c (kinds documented there)
cudaError_t generate_keys(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                          uint32_t p16, uint64_t key_mul, int64_t index_base, cudaStream_t stream);
cudaError_t generate_keys_total(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                                uint32_t p16, uint64_t key_mul, int64_t index_base, uint64_t n_total, cudaStream_t stream);

}  // namespace hj
