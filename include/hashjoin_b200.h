/*
 * hashjoin_b200.h — C ABI of libhashjoin_b200.so, the B200-native replacement for the build + probe path of
 * deveshv-99/mlir-HashJoin.  The library sits in the slot of the reference's shared_stuff/shared.so: it is loaded by
 * `mlir-cpu-runner --shared-libs=...` (run_test.sh:14,33) and its symbols are resolved by name from MLIR `func.call`s.
 *
 * Three groups of entry points (citations are file:line in the reference):
 *
 *  A. The six legacy helper symbols, bit-compatible with shared_stuff/shared.cpp:18-172 (expanded memref ABI,
 *     five scalars per rank-1 memref: allocated, aligned, offset, size, stride — shared.cpp:35, join_v1.ll:1262-1265).
 *
 *  B. The join entry points with the reference's own names and argument lists — the MLIR host wrappers
 *     @initializeHashTable / @buildTable / @countRows / @probeRelation (join_v1.mlir:54,77,110,149) — so the
 *     reference's @main (join_v1.mlir:525-649) runs unchanged once those four func.funcs are turned into
 *     `func.func private` declarations.  Each exists in the expanded ABI (plain name) and in the
 *     llvm.emit_c_interface ABI (`_mlir_ciface_<name>`, pointers to StridedMemRefType<T,1> descriptors).
 *
 *  C. The native B200 surface (`hashJoin*` for MLIR callers, `hj*` for C/C++/ctypes callers with explicit streams):
 *     opaque table/scratch workspaces sized by query, i32 and i64 keys, payload columns, radix partition, digest.
 *
 * All relation / table / scratch / result pointers are DEVICE pointers unless a name says Host.  The library uses the
 * CUDA runtime API on the current device (primary context), so pointers from mgpuMemAlloc / cudaMalloc / torch are
 * valid.  Entry points of groups A and B and the `hashJoin*` functions are synchronous on return; `hj*` functions
 * are asynchronous on the stream they are given unless documented otherwise.  Nothing throws across this boundary:
 * failures return a negative status (and print one line to stderr); hjLastErrorString() has the text (per calling thread).
 * Threads and devices: the `hj*` functions keep no per-call state on the host — layout decisions live in the device-resident table
 * header and scratch counters, small readbacks land in a pinned block owned by the calling thread — so different host threads may
 * drive different (table, scratch, stream) triples concurrently, on any device (the current device at the call is used). One table
 * may be probed from several streams at once with distinct scratch workspaces. The legacy group B entry points and hjJoinHost cache
 * their workspaces per process under a mutex: callable from any thread, serialised.
 * There is NO CPU fallback: without a CUDA device every join entry point fails with HJ_ERR_CUDA.
 */
#ifndef HASHJOIN_B200_H
#define HASHJOIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HJ_OK 0
#define HJ_ERR_ARG (-22)     /* bad argument: null pointer, non-unit stride, misaligned or too-small workspace */
#define HJ_ERR_CUDA (-5)     /* CUDA runtime error (no device, launch failure, out of memory) */
#define HJ_ERR_STATE (-2)    /* count / write on a table that was never built (or built for the other key width); hjJoinFused on a layout it cannot probe */

/* StridedMemRefType<T,1> as passed by llvm.emit_c_interface (Experiments/passing-memrefs.mlir:10, join_v1.ll:106-111) */
typedef struct { void* allocated; void* aligned; int64_t offset; int64_t sizes[1]; int64_t strides[1]; } HjMemRef1D;

/* ---- A. legacy helper symbols (shared_stuff/shared.cpp) -------------------------------------------------- */
void startTimer(void);                                                                   /* shared.cpp:18-20 */
void endTimer(void);                      /* prints "For <n>, time taken: <us> microseconds"  shared.cpp:23-31 */
void initRelationIndex(int32_t* allocated, int32_t* aligned, int64_t offset, int64_t size, int64_t stride); /* :35-41 */
void initRelationR(int32_t* allocated, int32_t* aligned, int64_t offset, int64_t size, int64_t stride);     /* :59-80 */
void initRelationS(int32_t* allocated, int32_t* aligned, int64_t offset, int64_t size, int64_t stride);     /* :83-116 */
/* HOST memrefs. 1 = result equals the nested-loop join, 0 = differs, -1 = more matches than result rows. :129-172 */
int32_t check(int32_t* rAlloc, int32_t* rAligned, int64_t rOff, int64_t rSize, int64_t rStride,
              int32_t* sAlloc, int32_t* sAligned, int64_t sOff, int64_t sSize, int64_t sStride,
              int32_t* orAlloc, int32_t* orAligned, int64_t orOff, int64_t orSize, int64_t orStride,
              int32_t* osAlloc, int32_t* osAligned, int64_t osOff, int64_t osSize, int64_t osStride);
/* Seeds for initRelationR / initRelationS (the reference's are unseeded, shared.cpp:62,86-87). Also read from the
 * environment variables HASHJOIN_SEED_R / HASHJOIN_SEED_S; unset = time-based like the reference. */
void hashJoinSetSeeds(uint64_t seedR, uint64_t seedS);

/* ---- B. the reference's join entry points, expanded ABI ---------------------------------------------------- */
#define HJ_MEMREF(T, n) T* n##Alloc, T* n##Aligned, int64_t n##Off, int64_t n##Size, int64_t n##Stride
/* join_v1.mlir:54   @initializeHashTable(index, memref<?xi32>) */
void initializeHashTable(int64_t hashTableSize, HJ_MEMREF(int32_t, head));
/* join_v1.mlir:77   @buildTable(memref<?xi32>, index, memref<?xi32>, memref<?xi32>, memref<?xindex>, memref<?xindex>, i32) */
void buildTable(HJ_MEMREF(int32_t, R), int64_t nR, HJ_MEMREF(int32_t, head), HJ_MEMREF(int32_t, lkey),
                HJ_MEMREF(int64_t, lrow), HJ_MEMREF(int64_t, lnext), int32_t hashTableSize);
/* join_v1.mlir:110  @countRows(...) -> index   (result size; the one mandatory host sync, :140-146) */
int64_t countRows(HJ_MEMREF(int32_t, S), int64_t nS, HJ_MEMREF(int32_t, head), HJ_MEMREF(int32_t, lkey),
                  HJ_MEMREF(int64_t, lrow), HJ_MEMREF(int64_t, lnext), HJ_MEMREF(int64_t, prefix), int32_t hashTableSize);
/* join_v1.mlir:149  @probeRelation(...) */
void probeRelation(HJ_MEMREF(int32_t, S), int64_t nS, int32_t hashTableSize, HJ_MEMREF(int32_t, head), HJ_MEMREF(int32_t, lkey),
                   HJ_MEMREF(int64_t, lrow), HJ_MEMREF(int64_t, lnext), HJ_MEMREF(int64_t, prefix),
                   HJ_MEMREF(int32_t, outR), HJ_MEMREF(int32_t, outS));
/* join_v1.mlir:43   @calculateNumberOfBlocks(index, index) -> index */
int64_t calculateNumberOfBlocks(int64_t totalThreads, int64_t threadsPerBlock);
/* releases every table the legacy surface cached (the reference never frees anything, join_v1.mlir:646) */
void hashJoinRelease(void);

/* same, llvm.emit_c_interface ABI */
void _mlir_ciface_initializeHashTable(int64_t hashTableSize, HjMemRef1D* head);
void _mlir_ciface_buildTable(HjMemRef1D* R, int64_t nR, HjMemRef1D* head, HjMemRef1D* lkey, HjMemRef1D* lrow, HjMemRef1D* lnext, int32_t hashTableSize);
int64_t _mlir_ciface_countRows(HjMemRef1D* S, int64_t nS, HjMemRef1D* head, HjMemRef1D* lkey, HjMemRef1D* lrow, HjMemRef1D* lnext,
                               HjMemRef1D* prefix, int32_t hashTableSize);
void _mlir_ciface_probeRelation(HjMemRef1D* S, int64_t nS, int32_t hashTableSize, HjMemRef1D* head, HjMemRef1D* lkey, HjMemRef1D* lrow,
                                HjMemRef1D* lnext, HjMemRef1D* prefix, HjMemRef1D* outR, HjMemRef1D* outS);
int32_t _mlir_ciface_check(HjMemRef1D* R, HjMemRef1D* S, HjMemRef1D* outR, HjMemRef1D* outS);
void _mlir_ciface_initRelationIndex(HjMemRef1D* a);
void _mlir_ciface_initRelationR(HjMemRef1D* a);
void _mlir_ciface_initRelationS(HjMemRef1D* a);

/* ---- C1. native MLIR surface: opaque workspaces (memref<?xi8>), sizes from the memref descriptors ---------- */
int64_t hashJoinTableBytes(int64_t nR);          /* i32 keys */
int64_t hashJoinTableBytesI64(int64_t nR);
int64_t hashJoinScratchBytes(int64_t nS);
int64_t hashJoinScratchBytesI64(int64_t nS);
int32_t hashJoinBuild(HJ_MEMREF(int32_t, R), HJ_MEMREF(int8_t, table));
int64_t hashJoinCount(HJ_MEMREF(int32_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch));
int32_t hashJoinWrite(HJ_MEMREF(int32_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch),
                      HJ_MEMREF(int32_t, outR), HJ_MEMREF(int32_t, outS));
int32_t hashJoinBuildI64(HJ_MEMREF(int64_t, R), HJ_MEMREF(int8_t, table));
int64_t hashJoinCountI64(HJ_MEMREF(int64_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch));
int32_t hashJoinWriteI64(HJ_MEMREF(int64_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch),
                         HJ_MEMREF(int32_t, outR), HJ_MEMREF(int32_t, outS));
/* late gather: out[k] = column[rowIds[k]] for one i32 payload column (what nested-loop.mlir:165-187 does per matched row) */
int32_t hashJoinGather(HJ_MEMREF(int32_t, column), HJ_MEMREF(int32_t, rowIds), HJ_MEMREF(int32_t, out));
int32_t _mlir_ciface_hashJoinGather(HjMemRef1D* column, HjMemRef1D* rowIds, HjMemRef1D* out);
int32_t _mlir_ciface_hashJoinBuild(HjMemRef1D* R, HjMemRef1D* table);
int64_t _mlir_ciface_hashJoinCount(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch);
int32_t _mlir_ciface_hashJoinWrite(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch, HjMemRef1D* outR, HjMemRef1D* outS);
int32_t _mlir_ciface_hashJoinBuildI64(HjMemRef1D* R, HjMemRef1D* table);
int64_t _mlir_ciface_hashJoinCountI64(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch);
int32_t _mlir_ciface_hashJoinWriteI64(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch, HjMemRef1D* outR, HjMemRef1D* outS);

/* ---- C2. native C surface: device pointers + explicit stream (cudaStream_t passed as void*) ---------------- */
int64_t hjTableBytes(int64_t nR, int32_t keyBytes);            /* keyBytes: 4 or 8 */
int64_t hjScratchBytes(int64_t nS, int32_t keyBytes);
/* K0+K1. payload == NULL: build row id = rowBase + i (join_v1.mlir:232). Row id 0xFFFFFFFF is reserved (EMPTY): rowBase + nR must
 * stay below it and payload values must not be 0xFFFFFFFF. dTable: 64-byte aligned (buckets are 32-byte vector loads paired into
 * 64-byte DRAM atoms). Asynchronous, except that a table beyond L2 reach reads its own header back once (one stream sync). */
int32_t hjBuild(const void* dR, int64_t nR, int32_t keyBytes, const uint32_t* dPayload, uint32_t rowBase,
                void* dTable, int64_t tableBytes, void* stream);
/* hjBuild with the table's policy stated per call instead of taken from the process defaults (hjSet*). The policy is stored in the
 * table header; every later hjCount / hjWrite on that table follows it, so tables built under different policies coexist.
 * policy = HJ_POLICY_DEFAULT or an OR of: HJ_POLICY_DENSE_* (one of), HJ_POLICY_RADIX, HJ_POLICY_LISTS_* (one of), HJ_POLICY_DUP_SAMPLE, HJ_POLICY_NO_SLICES. */
#define HJ_POLICY_DEFAULT 0xFFFFFFFFu
#define HJ_POLICY_DENSE_OFF 0u          /* never the direct-address layout */
#define HJ_POLICY_DENSE_CACHE 1u        /* direct-address layout for dense key ranges, lookups in the count pass (match cache) */
#define HJ_POLICY_DENSE_RANGE 2u        /* + gap-free unique ranges are counted by range test alone (library default) */
#define HJ_POLICY_RADIX (1u << 2)       /* tables beyond L2 reach (> 48 MB of buckets): radix join instead of a global hash table (default on) */
#define HJ_POLICY_LISTS_NEVER (0u << 3)
#define HJ_POLICY_LISTS_SAMPLED (1u << 3) /* hit lists for selective joins, decided on the device from a probe-key sample (default) */
#define HJ_POLICY_LISTS_ALWAYS (2u << 3)
#define HJ_POLICY_DUP_SAMPLE (1u << 5)  /* sample the build keys for duplicates before attempting a unique-key layout (default on) */
#define HJ_POLICY_NO_SLICES (1u << 7)   /* never the slice-ordered inline layout: every table beyond L2 reach takes the radix layout (default off) */
#define HJ_POLICY_TMA_COUNT (1u << 6)   /* experimental: TMA-staged streams in the direct-address count kernel (slower; default off) */
int32_t hjBuildEx(const void* dR, int64_t nR, int32_t keyBytes, const uint32_t* dPayload, uint32_t rowBase,
                  void* dTable, int64_t tableBytes, uint32_t policy, void* stream);
uint32_t hjDefaultPolicy(void);          /* the policy word hjBuild would use now */
/* K2+K3; hjCountResult copies the total back and synchronises the stream; hjCount does both. */
/* hjCountAsync queues the count; it first reads the table header back (256 bytes: one stream sync) so that only the kernels of the
 * layout actually built are launched. dScratch: 256-byte aligned. HJ_ERR_STATE when dTable holds no table for this key width. */
int32_t hjCountAsync(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes, void* stream);
int64_t hjCountResult(const void* dScratch, int64_t nS, int32_t keyBytes, void* stream);
int64_t hjCount(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes, void* stream);
/* K4. Writes exactly the pairs counted by the preceding hjCount on the same (S, table, scratch). probePayload == NULL:
 * probe row id = probeRowBase + j (join_v1.mlir:499-500 stores the thread index). Asynchronous. */
int32_t hjWrite(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, const void* dScratch,
                int32_t* dOutR, int32_t* dOutS, const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream);
/* hjCount / hjCountAsync for callers that already know the probe row ids at count time: the ids hjWrite would be given (payload
 * column, else probeRowBase + j). When the probe relation is radix-partitioned (tables beyond L2 reach) the copy then carries
 * these ids instead of the original index, and hjWrite needs no random gather through a payload column. hjWrite must be called
 * with the same dProbePayload / probeRowBase afterwards. */
int32_t hjCountAsyncRows(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                         const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream);
int64_t hjCountRows(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                    const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream);
/* K2+K3+K4 in ONE pass for callers that can bound the result size before probing (capacity = nS for a unique build): lookup,
 * decoupled look-back over the tiles' match counts, pairs streamed straight into the result columns; no match cache and no second
 * pass over the probe relation. The reference's call sequence (count, allocate, probe: join_v1.mlir:591,604-605) cannot use
 * it. Returns the number of pairs; pairs beyond `capacity` are counted but not written. HJ_ERR_STATE when the build keys were not
 * unique or the table is radix-partitioned (use hjCount + hjWrite). Synchronous. */
int64_t hjJoinFused(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                    int32_t* dOutR, int32_t* dOutS, int64_t capacity, const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream);
/* ---- the operators either side of the path ----
 * Semi-join (key-only mode): the probe rows that have at least one match, each exactly once, whatever the multiplicity of the build
 * keys. Same two-phase shape: hjSemiJoinCount returns the size, hjSemiJoinWrite fills dOutS (probe row ids: payload value, else
 * probeRowBase + j; state the same ones in both calls). Count-only: hjCount alone (no write pass is needed to know the size). */
int64_t hjSemiJoinCount(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                        const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream);
int32_t hjSemiJoinWrite(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, const void* dScratch, int32_t* dOutS,
                        const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream);
/* Late gather: dOut[k] = dColumn[dRowIds[k] - rowBase] for a payload column of 4- or 8-byte elements; dRowIds is one column of the
 * pair stream. Asynchronous. (nested-loop.mlir:165-187 copies matched rows inline; projectDescription.md:26 lists it as left out.) */
int32_t hjGather(const void* dColumn, int32_t elemBytes, const int32_t* dRowIds, int64_t n, uint32_t rowBase, void* dOut, void* stream);
/* Row materialisation in the reference's layout (nested-loop.mlir:165-187): result row k = the xCols columns of row dPairX[k] of
 * row-major table X, then columns 1 .. yCols-1 of row dPairY[k] of table Y (the key column is not stored twice). dResult: row-major
 * i32 [nPairs][xCols + yCols - 1]. hjExtractColumn copies column `col` of a row-major table into a contiguous key column. */
int32_t hjMaterializeRows(const int32_t* dTableX, int32_t xCols, const int32_t* dTableY, int32_t yCols, const int32_t* dPairX, const int32_t* dPairY,
                          int64_t nPairs, int32_t* dResult, void* stream);
int32_t hjExtractColumn(const int32_t* dTable, int64_t rows, int32_t cols, int32_t col, int32_t* dOut, void* stream);
/* Multi-column keys (projectDescription.md:28): two i32 key columns become one i64 key (a bijection), joined with keyBytes = 8. */
int32_t hjPackKeys2x32(const int32_t* dA, const int32_t* dB, int64_t n, int64_t* dOut, void* stream);
/* Non-integer keys (projectDescription.md:28): an f32 / f64 key column (elemBytes 4 / 8) becomes an i32 / i64 key column with IEEE
 * equality — equal numbers get equal keys (-0.0 joins +0.0) and a NaN joins nothing (probeSide != 0 marks the probe relation's column:
 * its NaNs get a pattern that differs from the build side's). Join the encoded columns with keyBytes = elemBytes. */
int32_t hjEncodeFloatKeys(const void* dColumn, int32_t elemBytes, int64_t n, int32_t probeSide, void* dOut, void* stream);
/* Selection (Experiments/selection.mlir:34-155): rows with `value OP constant`, count -> scan -> compacted write; output in input
 * order. dtype: 0 i32, 1 i64, 2 f32, 3 f64 (constant in iconst for integers, fconst for floats; comparisons are ordered: NaN fails).
 * op: 0 <, 1 <=, 2 >, 3 >=, 4 ==, 5 !=. hjSelectCount is synchronous and returns the size; hjSelectWrite fills dOutValues (same
 * dtype) and / or dOutRows (rowBase + i), either may be NULL. dScratch: hjSelectScratchBytes(n), 256-byte aligned. */
int64_t hjSelectScratchBytes(int64_t n);
int64_t hjSelectCount(const void* dColumn, int64_t n, int32_t dtype, int32_t op, int64_t iconst, double fconst, void* dScratch, int64_t scratchBytes, void* stream);
int32_t hjSelectWrite(const void* dColumn, int64_t n, int32_t dtype, int32_t op, int64_t iconst, double fconst, const void* dScratch,
                      void* dOutValues, int32_t* dOutRows, uint32_t rowBase, void* stream);
/* K5. Radix partition on the key hash into nParts (<= 256) contiguous ranges; dOffsets: u64[nParts+1]. Asynchronous. */
int64_t hjPartitionWorkspaceBytes(int64_t n, int32_t nParts);
int32_t hjPartition(const void* dKeys, const uint32_t* dRows, uint32_t rowBase, int64_t n, int32_t keyBytes, int32_t nParts,
                    void* dOutKeys, uint32_t* dOutRows, uint64_t* dOffsets, void* dWorkspace, int64_t workspaceBytes, void* stream);
/* K5 fused with the exchange (multi-GPU): hjPartitionCount fills dCounts[nParts] (tuples per partition) and keeps its per-CTA
 * count matrix in dWorkspace (hjPartitionWorkspaceBytes); after the ranks have all-gathered their counts, hjPartitionPush (same
 * keys, same workspace) stores every (key, row) straight into the receive buffer of the rank that owns its partition through
 * peer-mapped pointers (NVLink P2P): dPeerKeyPtrs / dPeerRowPtrs are device arrays of nParts buffer addresses, dCursors[p] is
 * the first element of this rank's region in partition p's buffer. The caller synchronises the ranks afterwards. Asynchronous. */
int32_t hjPartitionCount(const void* dKeys, int64_t n, int32_t keyBytes, int32_t nParts, uint64_t* dCounts, void* dWorkspace, int64_t workspaceBytes, void* stream);
int32_t hjPartitionPush(const void* dKeys, const uint32_t* dRows, uint32_t rowBase, int64_t n, int32_t keyBytes, int32_t nParts,
                        const uint64_t* dPeerKeyPtrs, const uint64_t* dPeerRowPtrs, const uint64_t* dCursors, void* dWorkspace, int64_t workspaceBytes, void* stream);
/* K6. Order-independent digest of a pair stream: hostOut2[0] = sum, hostOut2[1] = xor of mix64(r << 32 | s). Synchronous. */
int32_t hjPairDigest(const int32_t* dOutR, const int32_t* dOutS, int64_t n, uint64_t* hostOut2, void* stream);
/* Seeded device generators, bit-identical to the oracle's (kinds: 0 index, 1 unique, 2 uniform, 3 mixed, 4 fk, 5 zipf). */
int32_t hjGenerate(void* dOut, int64_t n, int32_t keyBytes, int32_t kind, uint64_t seed, int64_t lo, uint64_t domain,
                   uint32_t p16, uint64_t keyMul, int64_t indexBase, uint64_t nTotal, void* stream);
/* The keys of rows dRowIds[0 .. n) of the relation the same arguments describe (nTotal = its row count): lets a parity guard check
 * key equality of result pairs whose rows live on other ranks. */
int32_t hjGenerateAt(void* dOut, const uint32_t* dRowIds, int64_t n, int32_t keyBytes, int32_t kind, uint64_t seed, int64_t lo, uint64_t domain,
                     uint32_t p16, uint64_t keyMul, uint64_t nTotal, void* stream);
/* End-to-end convenience with HOST buffers (what the reference's @main does around the kernels, join_v1.mlir:558-615):
 * H2D of both relations, build, count, write, D2H of the pairs. Returns the result size; pairs are copied only when
 * hOutR/hOutS are non-NULL and capacity >= result size. Synchronous. */
int64_t hjJoinHost(const void* hR, int64_t nR, const void* hS, int64_t nS, int32_t keyBytes,
                   int32_t* hOutR, int32_t* hOutS, int64_t capacity);
/* hjJoinHost streams the probe relation through a ring of three device chunks and the pairs through two result slots, so the probe
 * side may be larger than GPU memory (projectDescription.md:23); the build side and its table must fit. Rows per chunk (default 2^24). */
void hjSetHostChunkRows(int64_t rows);
/* 0 forces the hash layout; 1: builds whose key range is at most 4x the row count use a direct-address table (with the match cache);
 * 2 (default): additionally a unique, gap-free key range (dense surrogate keys) is counted by range test alone and looked up once, in
 * the write pass (config 2: 1.84 instead of 1.99 ms). The policy in force at hjBuild decides. */
void hjSetAllowDense(int32_t on);
/* 1 (default): build relations whose hash table would not stay in L2 (> 48 MB of buckets) are partitioned first: up to 1 GB of buckets
 * (and unique keys) once, on the bucket hash, so that one global table is built and probed slice by slice; beyond that (or with
 * duplicate keys) twice, into <= 65 536 partitions joined in shared memory. 0 builds one hash table and probes in input order. */
void hjSetLocality(int32_t on);
/* 1: the direct-address count kernel moves its two streams with TMA bulk copies (cp.async.bulk, per-warp mbarriers); 0 (default): LDG/STG.
 * Experimental: measured 3.7x slower on config 2 (profiles/README.md). */
void hjSetTmaCount(int32_t on);
/* Hit lists for selective joins (unique build keys): the count pass appends (build row, probe position) per chunk instead of writing a
 * match-cache word per probe row, the write pass copies the lists. 1 (default): decided on the device from a sample of 2048 probe
 * keys (lists when < 35 % hit, relations of >= 2^20 probe rows); 0: never; 2: always. Same result either way. */
void hjSetSparse(int32_t policy);
/* Experiment switch: grid of BOTH direct-address probe kernels: 0 = one chunk per CTA, k > 0 = at most k resident waves striding over
 * the chunks. Default (not reachable through this call): count 2 waves, write one chunk per CTA — see hj_kernels.cu for the numbers. */
void hjSetDenseWaves(int32_t k);
/* 1 (default): builds of >= 2^18 rows look at 16 x 4 096 sampled rows for duplicate keys first and go straight to the grouped layout
 * when they find some (the inline, unique-key build is otherwise attempted and aborted); 0: always attempt the inline layout. */
void hjSetDupSample(int32_t on);
/* 1 (default): tables of 48 MB .. 1 GB of buckets with unique keys are built and probed as one hash table in table-slice order;
 * 0: they take the radix layout like the bigger ones (HJ_POLICY_NO_SLICES). */
void hjSetSliced(int32_t on);
/* Experiment switch: CTA shape of the radix-partition scatter kernel: 512 threads / 4 096-tuple tiles (default) or 256 / 2 048. */
void hjSetPartitionThreads(int32_t threads);
/* Layout the last hjBuild gave this table (diagnostic; one header readback): 0 = bucketised hash (unique keys), 1 = direct-address,
 * 2 = grouped (duplicate keys), 3 = radix-partitioned (beyond 1 GB of buckets, or duplicate keys beyond L2 reach); + 0x200 when layout 0 / 2
 * was built in table-slice order (48 MB .. 1 GB of buckets: both relations are partitioned once on the bucket hash); + 0x100 when the direct-address table is gap-free and unique, i.e. the count pass runs by range test. */
int32_t hjTableLayout(const void* dTable, void* stream);
/* Which probe path the last hjCount on this scratch took: 0 = match cache, 1 = hit lists (diagnostic; one 8-byte readback). */
int32_t hjProbePath(const void* dScratch, int64_t nS, int32_t keyBytes, void* stream);
const char* hjLastErrorString(void);
const char* hjVersion(void);

#ifdef __cplusplus
}
#endif
#endif /* HASHJOIN_B200_H */
