// hj_kernels.cuh — host-side launch interface of the sm_100a join kernels (hj_kernels.cu, hj_partition.cu, hj_radix.cu, hj_ops.cu).
// Everything here takes DEVICE pointers and an explicit stream. The only synchronising calls are the small readbacks the probe passes
// start with (count_rows_async, write_pairs: table header / scratch counters) and the big-table build's one look at its header.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hj {

struct TableHeader;
constexpr int HEADER_BYTES = 256;              // table workspace = header + body

// tile geometry shared by count/scan/write (must agree between the two probe passes)
constexpr int BLOCK_THREADS = 256;
constexpr int VECS_PER_THREAD = 2;
// consecutive tiles one CTA owns (the scan runs over chunks)
__host__ __device__ constexpr int chunk_tiles(int key_bytes) { return key_bytes == 4 ? 8 : 1; }
__host__ __device__ constexpr int keys_per_vec(int key_bytes) { return 16 / key_bytes; }
__host__ __device__ constexpr int tile_keys(int key_bytes) { return BLOCK_THREADS * VECS_PER_THREAD * keys_per_vec(key_bytes); }
__host__ __device__ constexpr int chunk_keys(int key_bytes) { return chunk_tiles(key_bytes) * tile_keys(key_bytes); }

// ---- policy word (TableHeader::policy; public names HJ_POLICY_* in include/hashjoin_b200.h) ----------------------------------------------
constexpr uint32_t POLICY_DENSE_MASK = 3u;        // 0 hash only, 1 direct-address + match cache, 2 + count by range test
constexpr uint32_t POLICY_RADIX = 1u << 2;        // tables beyond L2 reach: radix join (K5 x 2 + K7) instead of a hash table in global memory
constexpr int      POLICY_SPARSE_SHIFT = 3;       // 2 bits: hit lists never / sampled on the device / always
constexpr uint32_t POLICY_DUP_SAMPLE = 1u << 5;   // sample the build keys for duplicates before trying the unique-key layouts
constexpr uint32_t POLICY_TMA_COUNT = 1u << 6;    // experimental: TMA-staged streams in the direct-address count kernel
constexpr uint32_t POLICY_NO_SLICES = 1u << 7;    // never the slice-ordered inline layout: every table beyond L2 reach takes the radix layout
uint32_t default_policy();                        // what the hjSet* process defaults add up to

int64_t preferred_pairs(int64_t n_rows, int key_bytes);
int64_t table_bytes(int64_t n_rows, int key_bytes);
int64_t scratch_bytes(int64_t n_probe, int key_bytes);
int64_t num_chunks(int64_t n_probe, int key_bytes);
bool table_is_big(int64_t n_rows, int key_bytes);

// Scratch layout (device): [ match cache: u32 x round_up(n_probe, chunk) ][ as much again ][ offsets: u64 x (max(chunks, radix items) + 1) ]
// [ counters ][ scan block sums ][ warp counts ][ radix area: the partitioned copy of the probe relation, its offsets and work items ].
// Selective joins (few probe rows hit) use the first two areas together as per-chunk HIT LISTS instead: chunk c's hits are the
// first cnt[c] 8-byte entries (matched build row, position of the probe row inside the chunk) of its slice, see k_count_sparse.
constexpr int SCRATCH_COUNTERS = 16;
constexpr int CTR_TICKET_HASH = 0, CTR_TICKET_GROUP = 1, CTR_TICKET_GROUP_W = 2, CTR_SPARSE = 3 /* hit-list flag */, CTR_TICKET_SPARSE = 4,
              CTR_TICKET_RADIX = 5, CTR_CARRIED = 6 /* radix copy carries probe row ids, not indices */, CTR_TOTAL = 7 /* result size */, CTR_TICKET_RADIX_W = 8, CTR_SEMI = 9 /* semi-join: one result row per matching probe row */,
              CTR_TICKET_RADIX_E = 10, CTR_RADIX_MULTI = 11 /* radix items the match cache cannot describe */,
              CTR_SLICED = 12 /* the probe passes run over the slice-ordered copy of the probe relation */;
struct ScratchView {
  uint32_t* mcache;
  uint2* hit_list;
  uint32_t* run_start;                // grouped layout: first row-id slot of each probe row's run (the second half of the cache area)
  uint32_t* warp_counts;              // hit-list mode: entries in each of the 8 warp lists of a chunk
  unsigned long long* chunk_offsets;  // after scan: exclusive offsets per chunk (or per radix work item)
  int64_t nchunks;
  unsigned long long* counters;       // CTR_*
  unsigned long long* scan_sums;      // block totals of the two-level scan
  char* radix;                        // radix layout: partitioned probe relation, offsets, work items
};
ScratchView scratch_view(void* scratch, int64_t n_probe, int key_bytes);

void set_allow_dense(int on);
void set_locality(int on);
void set_sparse(int policy);
void set_dup_sample(int on);
void set_dense_waves(int k);    // experiment: grid of the direct-address probe kernels
void set_tma_count(int on);
void set_sliced(int on);
void set_partition_threads(int t);   // experiment: CTA shape of the partition scatter kernel (256 | 512)

cudaError_t readback(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t stream);   // <= 1 KB through the calling thread's pinned block; synchronises

// K0+K1: clear + build.  payload == nullptr -> row id = row_base + i  (join_v1.mlir:232 stores the thread index).
cudaError_t build_table(const void* R, int64_t nR, int key_bytes, const uint32_t* payload, uint32_t row_base,
                        void* table, int64_t table_bytes_, uint32_t policy, cudaStream_t stream);
// K2+K3: count + scan. The result size lands in counters[CTR_TOTAL]. carry_rows: the probe row ids (payload column or row base) are
// known now, so the radix layout's partitioned copy carries THEM instead of the original index; write_pairs then needs none.
cudaError_t count_rows_async(const void* S, int64_t nS, int key_bytes, const void* table, void* scratch,
                             bool carry_rows, const uint32_t* probe_payload, uint32_t probe_row_base, bool semi, cudaStream_t stream);
// K4: write pairs. outR == nullptr: only the probe rows are stored (semi-join).
cudaError_t write_pairs(const void* S, int64_t nS, int key_bytes, const void* table, const void* scratch,
                        int32_t* outR, int32_t* outS, const uint32_t* probe_payload, uint32_t probe_row_base, cudaStream_t stream);
// K3 alone: t[0 .. n) := exclusive prefix, t[n] := total (also *total_out when given)
void launch_scan(unsigned long long* t, int64_t n, unsigned long long* block_sums, unsigned long long* total_out, cudaStream_t stream);

// K2+K3+K4 fused (single pass, decoupled look-back): unique layouts only; the total lands where count_rows_async puts it.
cudaError_t join_fused_async(const void* S, int64_t nS, int key_bytes, const void* table, void* scratch, int32_t* outR, int32_t* outS, int64_t capacity,
                             const uint32_t* probe_payload, uint32_t probe_row_base, cudaStream_t stream);
cudaError_t read_table_mode(const void* table, uint32_t* mode, uint32_t* all_present, cudaStream_t stream);

// K5 (hj_partition.cu): radix partition of (key, row id) tuples. One pass on the owner hash (multi-GPU shuffle feed):
//   partition p occupies [offsets[p], offsets[p+1]) of out_keys / out_rows; offsets: u64[n_parts + 1] (device).
cudaError_t radix_partition(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int n_parts,
                            void* out_keys, uint32_t* out_rows, unsigned long long* offsets, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int64_t partition_workspace_bytes(int64_t n, int n_parts);
cudaError_t partition_count(const void* keys, int64_t n, int key_bytes, int n_parts, unsigned long long* counts, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
cudaError_t partition_push(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int n_parts,
                           void* const* peer_keys, uint32_t* const* peer_rows, const unsigned long long* cursors, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream);
// One pass on the top `bits` bits of the hash that picks the bucket pair of the inline (or, grouped = true, the grouped) table: slice s =
// the s-th 2^-bits of the table
int64_t slice_partition_workspace_bytes(int64_t n, int bits);
cudaError_t slice_partition(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int bits, bool grouped,
                            void* out_keys, uint32_t* out_rows, uint32_t* offsets, void* ws, int64_t ws_bytes, cudaStream_t stream);
// Two passes on the top bits1 + bits2 bits of radix_hash(key): 2^(bits1 + bits2) partitions, offsets u32[parts + 1] (device).
int64_t radix_partition2_workspace_bytes(int64_t n, int bits1, int bits2);
cudaError_t radix_partition2(const void* keys, const uint32_t* rows, uint32_t row_base, int64_t n, int key_bytes, int bits1, int bits2,
                             void* tmp_keys, uint32_t* tmp_rows, void* out_keys, uint32_t* out_rows, uint32_t* offsets, void* ws, int64_t ws_bytes, cudaStream_t stream);

// K7 (hj_radix.cu): radix join of tables beyond L2 reach
struct RjItem { uint32_t s0, s1, r0, r1; };    // probe tuples [s0, s1) of the partitioned probe copy against build tuples [r0, r1) of the table
void radix_bits(int64_t n_build, int* bits1, int* bits2);
int64_t radix_max_items(int64_t n_probe);
int64_t radix_table_bytes(int64_t n_build, int key_bytes);
int64_t radix_scratch_bytes(int64_t n_probe, int key_bytes);
// the slice-ordered copy of the probe relation lives in the same scratch area: [keys][row ids][offsets][translated row ids][partition workspace]
struct SliceArea { void* keys; uint32_t* rows; uint32_t* offsets; uint32_t* rows2; void* ws; int64_t ws_bytes; };
SliceArea slice_area(char* radix_area, int64_t n, int key_bytes);
cudaError_t radix_build(const void* R, int64_t nR, int key_bytes, const uint32_t* payload, uint32_t row_base, TableHeader* hdr, char* body, int64_t body_bytes, cudaStream_t stream);
cudaError_t radix_count(const void* S, int64_t nS, int key_bytes, const TableHeader& hdr_host, const char* body, char* scratch_area,
                        unsigned long long* item_totals, unsigned long long* scan_block_sums, unsigned long long* ticket, unsigned long long* total_out,
                        uint32_t* mcache, unsigned long long* n_multi, bool carry_rows, const uint32_t* probe_payload, uint32_t probe_row_base, bool semi, cudaStream_t stream);
cudaError_t radix_write(int64_t nS, int key_bytes, const TableHeader& hdr_host, const char* body, char* scratch_area, unsigned long long* item_offsets,
                        unsigned long long* ticket, unsigned long long* ticket_emit, const uint32_t* mcache, unsigned long long* n_multi,
                        int32_t* outR, int32_t* outS, bool carried_rows, const uint32_t* probe_payload, uint32_t probe_row_base, bool semi, cudaStream_t stream);

// hj_ops.cu: the operators either side of the join
cudaError_t gather_column(const void* column, int elem_bytes, const int32_t* rows, int64_t n, uint32_t row_base, void* out, cudaStream_t stream);
cudaError_t materialize_rows(const int32_t* tx, int x_cols, const int32_t* ty, int y_cols, const int32_t* pair_x, const int32_t* pair_y, int64_t n_pairs,
                             int32_t* result, cudaStream_t stream);
cudaError_t extract_column(const int32_t* table, int64_t rows, int cols, int col, int32_t* out, cudaStream_t stream);
cudaError_t pack_keys(const int32_t* a, const int32_t* b, int64_t n, long long* out, cudaStream_t stream);
cudaError_t encode_float_keys(const void* in, int64_t n, int elem_bytes, int probe_side, void* out, cudaStream_t stream);
int64_t select_scratch_bytes(int64_t n);
unsigned long long* select_total_ptr(void* scratch, int64_t n);
cudaError_t select_count(const void* col, int64_t n, int dtype, int op, long long iconst, double fconst, void* scratch, cudaStream_t stream);
cudaError_t select_write(const void* col, int64_t n, int dtype, int op, long long iconst, double fconst, const void* scratch, void* out_values, int32_t* out_rows, uint32_t row_base,
                         cudaStream_t stream);

// K6: verification helpers — order-independent digest of a pair stream: out[0] += sum(mix64(pair)), out[1] ^= xor.
cudaError_t pair_digest(const int32_t* outR, const int32_t* outS, int64_t n, unsigned long long* out2, cudaStream_t stream);

// seeded generators, bit-identical to oracle/oracle_join.c (kinds documented there)
cudaError_t generate_keys(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                          uint32_t p16, uint64_t key_mul, int64_t index_base, cudaStream_t stream);
cudaError_t generate_keys_total(void* out, int64_t n, int key_bytes, int kind, uint64_t seed, int64_t lo, uint64_t domain,
                                uint32_t p16, uint64_t key_mul, int64_t index_base, uint64_t n_total, const uint32_t* at, cudaStream_t stream);

}  // namespace hj
