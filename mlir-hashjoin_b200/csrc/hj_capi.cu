// hj_capi.cu — the C-ABI boundary (include/hashjoin_b200.h): legacy helper symbols, the reference's join entry
// points (expanded + llvm.emit_c_interface ABIs) and the native hj* surface.  No torch types, no C++ in signatures.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "../../include/hashjoin_b200.h"
#include "hj_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int32_t fail(int32_t code, const char* where, const char* what) {
  g_last_error = std::string(where) + ": " + what;
  fprintf(stderr, "[hashjoin_b200] %s\n", g_last_error.c_str());
  return code;
}
int32_t cuda_fail(const char* where, cudaError_t e) { return fail(HJ_ERR_CUDA, where, cudaGetErrorString(e)); }

#define HJ_CUDA(where, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(where, e_); } while (0)

inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool key_ok(int32_t kb) { return kb == 4 || kb == 8; }

// ---- legacy surface state: B200 tables cached per caller-visible head pointer -----------------------------
struct LegacyTable { void* table = nullptr; int64_t table_bytes = 0; void* scratch = nullptr; int64_t scratch_bytes = 0; bool built = false; };
std::mutex g_mu;
std::map<const void*, LegacyTable> g_tables;

// ---- hjJoinHost cache -------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr; int64_t bytes = 0;
  cudaError_t ensure(int64_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, (size_t)need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
DevBuf g_hR, g_hS, g_hT, g_hSc, g_hOr[2], g_hOs[2];

std::chrono::high_resolution_clock::time_point g_timer_start;
uint64_t g_seed_r = 0, g_seed_s = 0;
bool g_seed_r_set = false, g_seed_s_set = false;

inline uint64_t mix64h(uint64_t z) { z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z; }

void fill_random(int32_t* a, int64_t n, uint64_t seed) {
  // values in [1, 1e9] like shared.cpp:13-14,68 — but from a seedable counter-based generator
  const uint64_t range = 1000000000ULL;
  for (int64_t i = 0; i < n; i++) {
    uint64_t r = mix64h(seed * 0x9E3779B97F4A7C15ULL + mix64h((uint64_t)i + 0xD1B54A32D192ED03ULL));
    a[i] = (int32_t)(1 + (uint64_t)(((unsigned __int128)r * range) >> 64));
  }
}

template <typename T> inline T* mr_ptr(T* aligned, int64_t off) { return aligned ? aligned + off : nullptr; }
inline bool mr_ok(int64_t size, int64_t stride) { return size <= 1 || stride == 1; }

}  // namespace

extern "C" {

const char* hjLastErrorString(void) { return g_last_error.c_str(); }
const char* hjVersion(void) { return "hashjoin_b200 0.2 (sm_100a)"; }
void hjSetAllowDense(int32_t on) { hj::set_allow_dense(on); }
void hjSetLocality(int32_t on) { hj::set_locality(on); }
void hjSetTmaCount(int32_t on) { hj::set_tma_count(on); }
void hjSetSparse(int32_t policy) { hj::set_sparse(policy); }
void hjSetDenseWaves(int32_t k) { hj::set_dense_waves(k); }
void hjSetDupSample(int32_t on) { hj::set_dup_sample(on); }
void hjSetPartitionThreads(int32_t t) { hj::set_partition_threads(t); }
void hjSetSliced(int32_t on) { hj::set_sliced(on); }

// =========================================================================================================
// A. legacy helper symbols
// =========================================================================================================
void startTimer(void) { g_timer_start = std::chrono::high_resolution_clock::now(); }

void endTimer(void) {
  auto stop = std::chrono::high_resolution_clock::now();
  static int n = 0;
  long long us = std::chrono::duration_cast<std::chrono::microseconds>(stop - g_timer_start).count();
  printf("For %d, time taken: %lld microseconds\n", n, us);
  fflush(stdout);
  n++;
}

void hashJoinSetSeeds(uint64_t seedR, uint64_t seedS) { g_seed_r = seedR; g_seed_s = seedS; g_seed_r_set = g_seed_s_set = true; }

void initRelationIndex(int32_t*, int32_t* aligned, int64_t offset, int64_t size, int64_t) {
  int32_t* a = mr_ptr(aligned, offset);
  for (int64_t i = 0; i < size; i++) a[i] = (int32_t)i;
}
void initRelationR(int32_t*, int32_t* aligned, int64_t offset, int64_t size, int64_t) {
  uint64_t seed = g_seed_r;
  if (!g_seed_r_set) { const char* e = getenv("HASHJOIN_SEED_R"); seed = e ? strtoull(e, nullptr, 0) : (uint64_t)time(nullptr); }
  fill_random(mr_ptr(aligned, offset), size, seed);
}
void initRelationS(int32_t*, int32_t* aligned, int64_t offset, int64_t size, int64_t) {
  uint64_t seed = g_seed_s;
  if (!g_seed_s_set) { const char* e = getenv("HASHJOIN_SEED_S"); seed = e ? strtoull(e, nullptr, 0) : (uint64_t)std::random_device{}(); }
  fill_random(mr_ptr(aligned, offset), size, seed);
}

// Same verdicts as the reference's nested loop + sort + compare (shared.cpp:129-172), computed in
// O((n+m) log m + out log out): -1 when the true join has more pairs than the result holds (:158-160); otherwise the
// expected vector is the true pairs padded with (0,0) up to result_size (the reference value-initialises it, :140-141).
int32_t check(int32_t*, int32_t* rAligned, int64_t rOff, int64_t rSize, int64_t,
              int32_t*, int32_t* sAligned, int64_t sOff, int64_t sSize, int64_t,
              int32_t*, int32_t* orAligned, int64_t orOff, int64_t orSize, int64_t,
              int32_t*, int32_t* osAligned, int64_t osOff, int64_t osSize, int64_t) {
  if (orSize != osSize) { fail(HJ_ERR_ARG, "check", "result columns differ in length (the reference asserts here, shared.cpp:134)"); return 0; }
  const int32_t* R = mr_ptr(rAligned, rOff); const int32_t* S = mr_ptr(sAligned, sOff);
  const int32_t* oR = mr_ptr(orAligned, orOff); const int32_t* oS = mr_ptr(osAligned, osOff);
  const int64_t result_size = orSize;
  std::vector<std::pair<int32_t, int32_t>> skeys((size_t)sSize);           // (key, probe row) sorted by key then row
  for (int64_t j = 0; j < sSize; j++) skeys[(size_t)j] = {S[j], (int32_t)j};
  std::sort(skeys.begin(), skeys.end());
  std::vector<std::pair<int32_t, int32_t>> expect;
  expect.reserve((size_t)result_size);
  for (int64_t i = 0; i < rSize; i++) {
    auto lo = std::lower_bound(skeys.begin(), skeys.end(), std::make_pair(R[i], INT32_MIN));
    for (; lo != skeys.end() && lo->first == R[i]; ++lo) {
      if ((int64_t)expect.size() >= result_size) return -1;
      expect.emplace_back((int32_t)i, lo->second);
    }
  }
  expect.resize((size_t)result_size, std::make_pair(0, 0));
  std::vector<std::pair<int32_t, int32_t>> got((size_t)result_size);
  for (int64_t k = 0; k < result_size; k++) got[(size_t)k] = {oR[k], oS[k]};
  std::sort(expect.begin(), expect.end());
  std::sort(got.begin(), got.end());
  return expect == got ? 1 : 0;
}

// =========================================================================================================
// C2. native C surface
// =========================================================================================================
int64_t hjTableBytes(int64_t nR, int32_t keyBytes) { return (nR < 0 || !key_ok(keyBytes)) ? HJ_ERR_ARG : hj::table_bytes(nR, keyBytes); }
int64_t hjScratchBytes(int64_t nS, int32_t keyBytes) { return (nS < 0 || !key_ok(keyBytes)) ? HJ_ERR_ARG : hj::scratch_bytes(nS, keyBytes); }

static int32_t build_checked(const char* who, const void* dR, int64_t nR, int32_t keyBytes, const uint32_t* dPayload, uint32_t rowBase, void* dTable, int64_t tableBytes,
                             uint32_t policy, void* stream) {
  if (!key_ok(keyBytes) || nR < 0 || (nR > 0 && !dR) || !dTable) return fail(HJ_ERR_ARG, who, "null pointer or bad key width");
  if (nR > 0xFFFFFFFELL) return fail(HJ_ERR_ARG, who, "more than 2^32-2 build rows (row ids are 32-bit, join_v1.mlir:604)");
  if (!dPayload && (uint64_t)rowBase + (uint64_t)nR > 0xFFFFFFFFULL) return fail(HJ_ERR_ARG, who, "rowBase + nR reaches 0xFFFFFFFF, the EMPTY row id");
  // buckets are read with 32-byte vector loads and paired into 64-byte DRAM atoms: the workspace must be aligned to that
  if (reinterpret_cast<uintptr_t>(dTable) & 63) return fail(HJ_ERR_ARG, who, "table workspace must be 64-byte aligned");
  if (tableBytes < hj::table_bytes(nR, keyBytes)) return fail(HJ_ERR_ARG, who, "table workspace too small (see hjTableBytes)");
  HJ_CUDA(who, hj::build_table(dR, nR, keyBytes, dPayload, rowBase, dTable, tableBytes, policy, S_(stream)));
  return HJ_OK;
}
int32_t hjBuild(const void* dR, int64_t nR, int32_t keyBytes, const uint32_t* dPayload, uint32_t rowBase, void* dTable, int64_t tableBytes, void* stream) {
  return build_checked("hjBuild", dR, nR, keyBytes, dPayload, rowBase, dTable, tableBytes, hj::default_policy(), stream);
}
int32_t hjBuildEx(const void* dR, int64_t nR, int32_t keyBytes, const uint32_t* dPayload, uint32_t rowBase, void* dTable, int64_t tableBytes, uint32_t policy, void* stream) {
  if (policy == HJ_POLICY_DEFAULT) policy = hj::default_policy();
  else if ((policy & 3u) == 3u || ((policy >> 3) & 3u) == 3u || (policy >> 8)) return fail(HJ_ERR_ARG, "hjBuildEx", "unknown policy bits");
  return build_checked("hjBuildEx", dR, nR, keyBytes, dPayload, rowBase, dTable, tableBytes, policy, stream);
}
uint32_t hjDefaultPolicy(void) { return hj::default_policy(); }

static int32_t count_async(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                           bool carryRows, const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream, bool semi = false) {
  if (!key_ok(keyBytes) || nS < 0 || (nS > 0 && !dS) || !dTable || !dScratch) return fail(HJ_ERR_ARG, "hjCount", "null pointer or bad key width");
  if (nS > 0xFFFFFFFFLL) return fail(HJ_ERR_ARG, "hjCount", "more than 2^32-1 probe rows (row ids are 32-bit, join_v1.mlir:605)");
  if (carryRows && !dProbePayload && (uint64_t)probeRowBase + (uint64_t)nS > 0x100000000ULL) return fail(HJ_ERR_ARG, "hjCount", "probeRowBase + nS exceeds 2^32");
  if (reinterpret_cast<uintptr_t>(dTable) & 63) return fail(HJ_ERR_ARG, "hjCount", "table workspace must be 64-byte aligned");
  if (reinterpret_cast<uintptr_t>(dScratch) & 255) return fail(HJ_ERR_ARG, "hjCount", "scratch workspace must be 256-byte aligned");
  if (scratchBytes < hj::scratch_bytes(nS, keyBytes)) return fail(HJ_ERR_ARG, "hjCount", "scratch workspace too small (see hjScratchBytes)");
  cudaError_t e = hj::count_rows_async(dS, nS, keyBytes, dTable, dScratch, carryRows, dProbePayload, probeRowBase, semi, S_(stream));
  if (e == cudaErrorInvalidValue) return fail(HJ_ERR_STATE, "hjCount", "the table workspace holds no table built for this key width (call hjBuild first)");
  if (e != cudaSuccess) return cuda_fail("hjCount", e);
  return HJ_OK;
}

int32_t hjCountAsync(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes, void* stream) {
  return count_async(dS, nS, keyBytes, dTable, dScratch, scratchBytes, false, nullptr, 0, stream);
}
int32_t hjCountAsyncRows(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                         const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream) {
  return count_async(dS, nS, keyBytes, dTable, dScratch, scratchBytes, true, dProbePayload, probeRowBase, stream);
}
int64_t hjCountRows(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                    const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream) {
  int32_t rc = hjCountAsyncRows(dS, nS, keyBytes, dTable, dScratch, scratchBytes, dProbePayload, probeRowBase, stream);
  if (rc != HJ_OK) return rc;
  return hjCountResult(dScratch, nS, keyBytes, stream);
}

int64_t hjCountResult(const void* dScratch, int64_t nS, int32_t keyBytes, void* stream) {
  if (!dScratch || !key_ok(keyBytes) || nS < 0) return fail(HJ_ERR_ARG, "hjCountResult", "bad argument");
  hj::ScratchView sv = hj::scratch_view(const_cast<void*>(dScratch), nS, keyBytes);
  unsigned long long total = 0;
  HJ_CUDA("hjCountResult", hj::readback(&total, sv.counters + hj::CTR_TOTAL, 8, S_(stream)));
  return (int64_t)total;
}

int32_t hjTableLayout(const void* dTable, void* stream) {
  if (!dTable) return fail(HJ_ERR_ARG, "hjTableLayout", "null table");
  uint32_t mode = 0, all_present = 0;
  HJ_CUDA("hjTableLayout", hj::read_table_mode(dTable, &mode, &all_present, S_(stream)));
  return (int32_t)(mode | (all_present ? 0x100u : 0u));
}

int32_t hjProbePath(const void* dScratch, int64_t nS, int32_t keyBytes, void* stream) {
  if (!dScratch || !key_ok(keyBytes) || nS < 0) return fail(HJ_ERR_ARG, "hjProbePath", "bad argument");
  hj::ScratchView sv = hj::scratch_view(const_cast<void*>(dScratch), nS, keyBytes);
  unsigned long long flag = 0;
  HJ_CUDA("hjProbePath", hj::readback(&flag, sv.counters + hj::CTR_SPARSE, 8, S_(stream)));
  return flag ? 1 : 0;
}

int64_t hjCount(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes, void* stream) {
  int32_t rc = hjCountAsync(dS, nS, keyBytes, dTable, dScratch, scratchBytes, stream);
  if (rc != HJ_OK) return rc;
  return hjCountResult(dScratch, nS, keyBytes, stream);
}

int32_t hjWrite(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, const void* dScratch,
                int32_t* dOutR, int32_t* dOutS, const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream) {
  if (!key_ok(keyBytes) || nS < 0 || (nS > 0 && !dS) || !dTable || !dScratch) return fail(HJ_ERR_ARG, "hjWrite", "null pointer or bad key width");
  if ((reinterpret_cast<uintptr_t>(dTable) & 63) || (reinterpret_cast<uintptr_t>(dScratch) & 255)) return fail(HJ_ERR_ARG, "hjWrite", "misaligned table or scratch workspace");
  cudaError_t e = hj::write_pairs(dS, nS, keyBytes, dTable, dScratch, dOutR, dOutS, dProbePayload, probeRowBase, S_(stream));
  if (e == cudaErrorInvalidValue) return fail(HJ_ERR_STATE, "hjWrite", "the table workspace holds no table built for this key width");
  if (e != cudaSuccess) return cuda_fail("hjWrite", e);
  return HJ_OK;
}

// ---- semi-join (key-only mode): the probe rows that have at least one match, each once ------------------------------------
int64_t hjSemiJoinCount(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                        const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream) {
  int32_t rc = count_async(dS, nS, keyBytes, dTable, dScratch, scratchBytes, true, dProbePayload, probeRowBase, stream, true);
  if (rc != HJ_OK) return rc;
  return hjCountResult(dScratch, nS, keyBytes, stream);
}
int32_t hjSemiJoinWrite(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, const void* dScratch, int32_t* dOutS,
                        const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream) {
  return hjWrite(dS, nS, keyBytes, dTable, dScratch, nullptr, dOutS, dProbePayload, probeRowBase, stream);
}

// ---- late gather / row materialisation (nested-loop.mlir:165-187) ------------------------------------------------------------
int32_t hjGather(const void* dColumn, int32_t elemBytes, const int32_t* dRowIds, int64_t n, uint32_t rowBase, void* dOut, void* stream) {
  if (n < 0 || (elemBytes != 4 && elemBytes != 8) || (n > 0 && (!dColumn || !dRowIds || !dOut))) return fail(HJ_ERR_ARG, "hjGather", "null pointer or element width not 4 / 8");
  HJ_CUDA("hjGather", hj::gather_column(dColumn, elemBytes, dRowIds, n, rowBase, dOut, S_(stream)));
  return HJ_OK;
}
int32_t hjMaterializeRows(const int32_t* dTableX, int32_t xCols, const int32_t* dTableY, int32_t yCols, const int32_t* dPairX, const int32_t* dPairY, int64_t nPairs,
                          int32_t* dResult, void* stream) {
  if (nPairs < 0 || xCols < 1 || yCols < 1 || (nPairs > 0 && (!dTableX || !dTableY || !dPairX || !dPairY || !dResult))) return fail(HJ_ERR_ARG, "hjMaterializeRows", "bad argument");
  HJ_CUDA("hjMaterializeRows", hj::materialize_rows(dTableX, xCols, dTableY, yCols, dPairX, dPairY, nPairs, dResult, S_(stream)));
  return HJ_OK;
}
int32_t hjExtractColumn(const int32_t* dTable, int64_t rows, int32_t cols, int32_t col, int32_t* dOut, void* stream) {
  if (rows < 0 || cols < 1 || col < 0 || col >= cols || (rows > 0 && (!dTable || !dOut))) return fail(HJ_ERR_ARG, "hjExtractColumn", "bad argument");
  HJ_CUDA("hjExtractColumn", hj::extract_column(dTable, rows, cols, col, dOut, S_(stream)));
  return HJ_OK;
}
int32_t hjPackKeys2x32(const int32_t* dA, const int32_t* dB, int64_t n, int64_t* dOut, void* stream) {
  if (n < 0 || (n > 0 && (!dA || !dB || !dOut))) return fail(HJ_ERR_ARG, "hjPackKeys2x32", "bad argument");
  HJ_CUDA("hjPackKeys2x32", hj::pack_keys(dA, dB, n, reinterpret_cast<long long*>(dOut), S_(stream)));
  return HJ_OK;
}

int32_t hjEncodeFloatKeys(const void* dColumn, int32_t elemBytes, int64_t n, int32_t probeSide, void* dOut, void* stream) {
  if (n < 0 || (elemBytes != 4 && elemBytes != 8) || (n > 0 && (!dColumn || !dOut))) return fail(HJ_ERR_ARG, "hjEncodeFloatKeys", "bad argument (elemBytes is 4 for f32 -> i32 keys or 8 for f64 -> i64 keys)");
  HJ_CUDA("hjEncodeFloatKeys", hj::encode_float_keys(dColumn, n, elemBytes, probeSide ? 1 : 0, dOut, S_(stream)));
  return HJ_OK;
}

// ---- selection (Experiments/selection.mlir:34-155): count -> scan -> write ----------------------------------------------------
int64_t hjSelectScratchBytes(int64_t n) { return n < 0 ? HJ_ERR_ARG : hj::select_scratch_bytes(n); }
static bool select_args_ok(const void* col, int64_t n, int32_t dtype, int32_t op, const void* scratch) {
  return n >= 0 && dtype >= 0 && dtype <= 3 && op >= 0 && op <= 5 && scratch && (n == 0 || col) && (reinterpret_cast<uintptr_t>(scratch) & 255) == 0;
}
int64_t hjSelectCount(const void* dColumn, int64_t n, int32_t dtype, int32_t op, int64_t iconst, double fconst, void* dScratch, int64_t scratchBytes, void* stream) {
  if (!select_args_ok(dColumn, n, dtype, op, dScratch) || scratchBytes < hj::select_scratch_bytes(n)) return fail(HJ_ERR_ARG, "hjSelectCount", "bad argument, or scratch misaligned / too small (see hjSelectScratchBytes)");
  if (n > 0xFFFFFFFFLL) return fail(HJ_ERR_ARG, "hjSelectCount", "more than 2^32-1 rows (row ids are 32-bit)");
  HJ_CUDA("hjSelectCount", hj::select_count(dColumn, n, dtype, op, iconst, fconst, dScratch, S_(stream)));
  unsigned long long total = 0;
  HJ_CUDA("hjSelectCount", hj::readback(&total, hj::select_total_ptr(dScratch, n), 8, S_(stream)));
  return (int64_t)total;
}
int32_t hjSelectWrite(const void* dColumn, int64_t n, int32_t dtype, int32_t op, int64_t iconst, double fconst, const void* dScratch,
                      void* dOutValues, int32_t* dOutRows, uint32_t rowBase, void* stream) {
  if (!select_args_ok(dColumn, n, dtype, op, dScratch)) return fail(HJ_ERR_ARG, "hjSelectWrite", "bad argument or misaligned scratch");
  HJ_CUDA("hjSelectWrite", hj::select_write(dColumn, n, dtype, op, iconst, fconst, dScratch, dOutValues, dOutRows, rowBase, S_(stream)));
  return HJ_OK;
}

// Single pass (K2+K3+K4 fused, decoupled look-back). Synchronous: returns the number of pairs found; pairs beyond `capacity`
// are counted, not written (the caller then retries with count + write or a bigger result). Grouped tables (duplicate build
// keys) are not handled by the fused kernel: HJ_ERR_STATE, use hjCount + hjWrite.
int64_t hjJoinFused(const void* dS, int64_t nS, int32_t keyBytes, const void* dTable, void* dScratch, int64_t scratchBytes,
                    int32_t* dOutR, int32_t* dOutS, int64_t capacity, const uint32_t* dProbePayload, uint32_t probeRowBase, void* stream) {
  if (!key_ok(keyBytes) || nS < 0 || (nS > 0 && !dS) || !dTable || !dScratch || capacity < 0 || (capacity > 0 && (!dOutR || !dOutS)))
    return fail(HJ_ERR_ARG, "hjJoinFused", "null pointer or bad key width");
  if (nS > 0xFFFFFFFFLL) return fail(HJ_ERR_ARG, "hjJoinFused", "more than 2^32-1 probe rows (row ids are 32-bit, join_v1.mlir:605)");
  if ((reinterpret_cast<uintptr_t>(dScratch) & 255) || (reinterpret_cast<uintptr_t>(dTable) & 63) || scratchBytes < hj::scratch_bytes(nS, keyBytes))
    return fail(HJ_ERR_ARG, "hjJoinFused", "workspace misaligned or scratch too small (see hjScratchBytes)");
  HJ_CUDA("hjJoinFused", hj::join_fused_async(dS, nS, keyBytes, dTable, dScratch, dOutR, dOutS, capacity, dProbePayload, probeRowBase, S_(stream)));
  uint32_t mode = 0;
  HJ_CUDA("hjJoinFused", hj::read_table_mode(dTable, &mode, nullptr, S_(stream)));
  if ((mode & 0xFF) >= 2) return fail(HJ_ERR_STATE, "hjJoinFused", "grouped (duplicate build keys) or radix (beyond L2 reach) table: use hjCount + hjWrite");
  return hjCountResult(dScratch, nS, keyBytes, stream);
}

int64_t hjPartitionWorkspaceBytes(int64_t n, int32_t nParts) { return hj::partition_workspace_bytes(n, nParts); }

int32_t hjPartition(const void* dKeys, const uint32_t* dRows, uint32_t rowBase, int64_t n, int32_t keyBytes, int32_t nParts,
                    void* dOutKeys, uint32_t* dOutRows, uint64_t* dOffsets, void* dWorkspace, int64_t workspaceBytes, void* stream) {
  if (!key_ok(keyBytes) || n < 0 || (n > 0 && (!dKeys || !dOutKeys || !dOutRows)) || !dOffsets || !dWorkspace)
    return fail(HJ_ERR_ARG, "hjPartition", "null pointer or bad key width");
  HJ_CUDA("hjPartition", hj::radix_partition(dKeys, dRows, rowBase, n, keyBytes, nParts, dOutKeys, dOutRows,
                                             reinterpret_cast<unsigned long long*>(dOffsets), dWorkspace, workspaceBytes, S_(stream)));
  return HJ_OK;
}

int32_t hjPartitionCount(const void* dKeys, int64_t n, int32_t keyBytes, int32_t nParts, uint64_t* dCounts, void* dWorkspace, int64_t workspaceBytes, void* stream) {
  if (!key_ok(keyBytes) || n < 0 || (n > 0 && !dKeys) || !dCounts || !dWorkspace) return fail(HJ_ERR_ARG, "hjPartitionCount", "null pointer or bad key width");
  HJ_CUDA("hjPartitionCount", hj::partition_count(dKeys, n, keyBytes, nParts, reinterpret_cast<unsigned long long*>(dCounts), dWorkspace, workspaceBytes, S_(stream)));
  return HJ_OK;
}

int32_t hjPartitionPush(const void* dKeys, const uint32_t* dRows, uint32_t rowBase, int64_t n, int32_t keyBytes, int32_t nParts,
                        const uint64_t* dPeerKeyPtrs, const uint64_t* dPeerRowPtrs, const uint64_t* dCursors, void* dWorkspace, int64_t workspaceBytes, void* stream) {
  if (!key_ok(keyBytes) || n < 0 || (n > 0 && !dKeys) || !dPeerKeyPtrs || !dPeerRowPtrs || !dCursors || !dWorkspace) return fail(HJ_ERR_ARG, "hjPartitionPush", "null pointer or bad key width");
  HJ_CUDA("hjPartitionPush", hj::partition_push(dKeys, dRows, rowBase, n, keyBytes, nParts, reinterpret_cast<void* const*>(dPeerKeyPtrs),
                                                reinterpret_cast<uint32_t* const*>(dPeerRowPtrs), reinterpret_cast<const unsigned long long*>(dCursors),
                                                dWorkspace, workspaceBytes, S_(stream)));
  return HJ_OK;
}

int32_t hjPairDigest(const int32_t* dOutR, const int32_t* dOutS, int64_t n, uint64_t* hostOut2, void* stream) {
  if (!hostOut2 || n < 0 || (n > 0 && (!dOutR || !dOutS))) return fail(HJ_ERR_ARG, "hjPairDigest", "bad argument");
  unsigned long long* d = nullptr;                                 // per call (stream-ordered): no buffer shared between threads or devices
  HJ_CUDA("hjPairDigest", cudaMallocAsync(&d, 16, S_(stream)));
  cudaError_t e = hj::pair_digest(dOutR, dOutS, n, d, S_(stream));
  if (e == cudaSuccess) e = hj::readback(hostOut2, d, 16, S_(stream));
  cudaFreeAsync(d, S_(stream));
  if (e != cudaSuccess) return cuda_fail("hjPairDigest", e);
  return HJ_OK;
}

int32_t hjGenerate(void* dOut, int64_t n, int32_t keyBytes, int32_t kind, uint64_t seed, int64_t lo, uint64_t domain,
                   uint32_t p16, uint64_t keyMul, int64_t indexBase, uint64_t nTotal, void* stream) {
  if (!key_ok(keyBytes) || n < 0 || (n > 0 && !dOut) || kind < 0 || kind > 5) return fail(HJ_ERR_ARG, "hjGenerate", "bad argument");
  if (nTotal == 0) nTotal = (uint64_t)(indexBase + n);
  HJ_CUDA("hjGenerate", hj::generate_keys_total(dOut, n, keyBytes, kind, seed, lo, domain, p16, keyMul, indexBase, nTotal, nullptr, S_(stream)));
  return HJ_OK;
}
int32_t hjGenerateAt(void* dOut, const uint32_t* dRowIds, int64_t n, int32_t keyBytes, int32_t kind, uint64_t seed, int64_t lo, uint64_t domain,
                     uint32_t p16, uint64_t keyMul, uint64_t nTotal, void* stream) {
  if (!key_ok(keyBytes) || n < 0 || (n > 0 && (!dOut || !dRowIds)) || kind < 0 || kind > 5 || nTotal == 0) return fail(HJ_ERR_ARG, "hjGenerateAt", "bad argument");
  HJ_CUDA("hjGenerateAt", hj::generate_keys_total(dOut, n, keyBytes, kind, seed, lo, domain, p16, keyMul, 0, nTotal, dRowIds, S_(stream)));
  return HJ_OK;
}

// Host buffers in, host pairs out. The probe relation moves in chunks so that the three PCIe/compute legs overlap:
//   stream IN : H2D of probe chunks i+1, i+2 into a ring of THREE device slots (the probe relation never sits on the device whole:
//               it may be larger than GPU memory — projectDescription.md:23 "relations that don't fit on GPU"; the build side must fit)
//   stream CMP: count(i) -> result-size readback -> write(i) into one of TWO result slots
//   stream OUT: D2H of the pairs of chunk i           (PCIe is full duplex: runs against the H2D of later chunks)
// Pairs are appended chunk by chunk (probe_row = chunk base + local row), which is a valid order for a multiset result.
// Device footprint: build column + table + 3 probe chunks + scratch for one chunk + 2 result slots of the largest chunk result.
static int64_t g_host_chunk_rows = (int64_t)1 << 24;
void hjSetHostChunkRows(int64_t rows) { if (rows >= 1024) g_host_chunk_rows = rows; }

int64_t hjJoinHost(const void* hR, int64_t nR, const void* hS, int64_t nS, int32_t keyBytes, int32_t* hOutR, int32_t* hOutS, int64_t capacity) {
  if (!key_ok(keyBytes) || nR < 0 || nS < 0 || (nR > 0 && !hR) || (nS > 0 && !hS)) return fail(HJ_ERR_ARG, "hjJoinHost", "bad argument");
  if (nS > 0xFFFFFFFFLL) return fail(HJ_ERR_ARG, "hjJoinHost", "more than 2^32-1 probe rows (row ids are 32-bit, join_v1.mlir:605)");
  std::lock_guard<std::mutex> lk(g_mu);
  const int64_t chunk_rows = g_host_chunk_rows;
  const int64_t nch = (nS + chunk_rows - 1) / chunk_rows;
  constexpr int RING = 3, OUT_RING = 2;
  const bool want_out = hOutR && hOutS && capacity > 0;
  const int64_t tb = hj::table_bytes(nR, keyBytes), sb = hj::scratch_bytes(std::min(nS, chunk_rows), keyBytes);
  const int64_t slot_bytes = (std::min(nS, chunk_rows) * keyBytes + 255) / 256 * 256;
  HJ_CUDA("hjJoinHost", g_hR.ensure(std::max<int64_t>(nR * keyBytes, 16)));
  HJ_CUDA("hjJoinHost", g_hS.ensure(std::max<int64_t>(slot_bytes * std::min<int64_t>(RING, std::max<int64_t>(nch, 1)), 16)));
  HJ_CUDA("hjJoinHost", g_hT.ensure(tb));
  HJ_CUDA("hjJoinHost", g_hSc.ensure(sb));
  static cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  static cudaEvent_t ev_in[RING], ev_w = nullptr, ev_out[OUT_RING];
  if (!s_in) {
    HJ_CUDA("hjJoinHost", cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    HJ_CUDA("hjJoinHost", cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
    HJ_CUDA("hjJoinHost", cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    HJ_CUDA("hjJoinHost", cudaEventCreateWithFlags(&ev_w, cudaEventDisableTiming));
    for (auto& e : ev_in) HJ_CUDA("hjJoinHost", cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ev_out) HJ_CUDA("hjJoinHost", cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const char* hSb = reinterpret_cast<const char*>(hS);
  char* dSb = reinterpret_cast<char*>(g_hS.p);
  auto issue_h2d = [&](int64_t i) -> cudaError_t {                                     // chunk i -> ring slot i % RING
    const int64_t off = i * chunk_rows, n = std::min(chunk_rows, nS - off);
    cudaError_t e = cudaMemcpyAsync(dSb + (i % RING) * slot_bytes, hSb + off * keyBytes, (size_t)n * keyBytes, cudaMemcpyHostToDevice, s_in);
    return e == cudaSuccess ? cudaEventRecord(ev_in[i % RING], s_in) : e;
  };
  if (nR) HJ_CUDA("hjJoinHost", cudaMemcpyAsync(g_hR.p, hR, (size_t)nR * keyBytes, cudaMemcpyHostToDevice, s_cmp));   // join_v1.mlir:558-561
  int32_t rc = hjBuild(g_hR.p, nR, keyBytes, nullptr, 0, g_hT.p, tb, s_cmp);
  if (rc != HJ_OK) return rc;
  for (int64_t i = 0; i < std::min<int64_t>(nch, RING - 1); i++) HJ_CUDA("hjJoinHost", issue_h2d(i));
  int64_t total = 0;
  bool overflow = !want_out;
  for (int64_t i = 0; i < nch; i++) {
    const int64_t off = i * chunk_rows, n = std::min(chunk_rows, nS - off);
    const char* dChunk = dSb + (i % RING) * slot_bytes;
    HJ_CUDA("hjJoinHost", cudaStreamWaitEvent(s_cmp, ev_in[i % RING], 0));
    const int64_t c = hjCount(dChunk, n, keyBytes, g_hT.p, g_hSc.p, sb, s_cmp);                                       // :591 (per chunk); synchronises s_cmp
    if (c < 0) return c;
    // s_cmp is idle now, so write(i-1) has finished reading slot (i-1) % RING == (i+2) % RING: refill it
    if (i + RING - 1 < nch) HJ_CUDA("hjJoinHost", issue_h2d(i + RING - 1));
    if (!overflow && total + c > capacity) overflow = true;
    if (!overflow && c > 0) {                                                                                        // :600-615
      DevBuf& oR = g_hOr[i % OUT_RING]; DevBuf& oS = g_hOs[i % OUT_RING];
      if (oR.bytes < c * 4 || oS.bytes < c * 4) {                      // growing a slot frees it: its last D2H must be done
        HJ_CUDA("hjJoinHost", cudaEventSynchronize(ev_out[i % OUT_RING]));
        HJ_CUDA("hjJoinHost", oR.ensure(c * 4)); HJ_CUDA("hjJoinHost", oS.ensure(c * 4));
      }
      HJ_CUDA("hjJoinHost", cudaStreamWaitEvent(s_cmp, ev_out[i % OUT_RING], 0));       // the slot's previous pairs have left the device
      int32_t* dR = reinterpret_cast<int32_t*>(oR.p); int32_t* dS = reinterpret_cast<int32_t*>(oS.p);
      rc = hjWrite(dChunk, n, keyBytes, g_hT.p, g_hSc.p, dR, dS, nullptr, (uint32_t)off, s_cmp);
      if (rc != HJ_OK) return rc;
      HJ_CUDA("hjJoinHost", cudaEventRecord(ev_w, s_cmp));
      HJ_CUDA("hjJoinHost", cudaStreamWaitEvent(s_out, ev_w, 0));
      HJ_CUDA("hjJoinHost", cudaMemcpyAsync(hOutR + total, dR, (size_t)c * 4, cudaMemcpyDeviceToHost, s_out));
      HJ_CUDA("hjJoinHost", cudaMemcpyAsync(hOutS + total, dS, (size_t)c * 4, cudaMemcpyDeviceToHost, s_out));
      HJ_CUDA("hjJoinHost", cudaEventRecord(ev_out[i % OUT_RING], s_out));
    }
    total += c;
  }
  HJ_CUDA("hjJoinHost", cudaStreamSynchronize(s_in));
  HJ_CUDA("hjJoinHost", cudaStreamSynchronize(s_cmp));
  HJ_CUDA("hjJoinHost", cudaStreamSynchronize(s_out));
  return total;
}

// =========================================================================================================
// C1. native MLIR surface (expanded ABI); synchronous on return
// =========================================================================================================
int64_t hashJoinTableBytes(int64_t nR) { return hjTableBytes(nR, 4); }
int64_t hashJoinTableBytesI64(int64_t nR) { return hjTableBytes(nR, 8); }
int64_t hashJoinScratchBytes(int64_t nS) { return hjScratchBytes(nS, 4); }
int64_t hashJoinScratchBytesI64(int64_t nS) { return hjScratchBytes(nS, 8); }

#define MR_ARGS_OK(n) mr_ok(n##Size, n##Stride)

static int32_t mlir_build(const void* R, int64_t nR, bool ok, int32_t kb, int8_t* table, int64_t tableBytes) {
  if (!ok) return fail(HJ_ERR_ARG, "hashJoinBuild", "non-unit stride memref");
  int32_t rc = hjBuild(R, nR, kb, nullptr, 0, table, tableBytes, nullptr);
  if (rc != HJ_OK) return rc;
  cudaError_t e = cudaStreamSynchronize(nullptr);
  return e == cudaSuccess ? HJ_OK : cuda_fail("hashJoinBuild", e);
}
static int32_t mlir_write(const void* S, int64_t nS, bool ok, int32_t kb, const int8_t* table, const int8_t* scratch, int32_t* outR, int32_t* outS) {
  if (!ok) return fail(HJ_ERR_ARG, "hashJoinWrite", "non-unit stride memref");
  int32_t rc = hjWrite(S, nS, kb, table, scratch, outR, outS, nullptr, 0, nullptr);
  if (rc != HJ_OK) return rc;
  cudaError_t e = cudaStreamSynchronize(nullptr);
  return e == cudaSuccess ? HJ_OK : cuda_fail("hashJoinWrite", e);
}

int32_t hashJoinBuild(HJ_MEMREF(int32_t, R), HJ_MEMREF(int8_t, table)) {
  (void)RAlloc; (void)tableAlloc;
  return mlir_build(mr_ptr(RAligned, ROff), RSize, MR_ARGS_OK(R) && MR_ARGS_OK(table), 4, mr_ptr(tableAligned, tableOff), tableSize);
}
int32_t hashJoinBuildI64(HJ_MEMREF(int64_t, R), HJ_MEMREF(int8_t, table)) {
  (void)RAlloc; (void)tableAlloc;
  return mlir_build(mr_ptr(RAligned, ROff), RSize, MR_ARGS_OK(R) && MR_ARGS_OK(table), 8, mr_ptr(tableAligned, tableOff), tableSize);
}
int64_t hashJoinCount(HJ_MEMREF(int32_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch)) {
  (void)SAlloc; (void)tableAlloc; (void)scratchAlloc; (void)tableSize;
  if (!(MR_ARGS_OK(S) && MR_ARGS_OK(table) && MR_ARGS_OK(scratch))) return fail(HJ_ERR_ARG, "hashJoinCount", "non-unit stride memref");
  return hjCount(mr_ptr(SAligned, SOff), SSize, 4, mr_ptr(tableAligned, tableOff), mr_ptr(scratchAligned, scratchOff), scratchSize, nullptr);
}
int64_t hashJoinCountI64(HJ_MEMREF(int64_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch)) {
  (void)SAlloc; (void)tableAlloc; (void)scratchAlloc; (void)tableSize;
  if (!(MR_ARGS_OK(S) && MR_ARGS_OK(table) && MR_ARGS_OK(scratch))) return fail(HJ_ERR_ARG, "hashJoinCountI64", "non-unit stride memref");
  return hjCount(mr_ptr(SAligned, SOff), SSize, 8, mr_ptr(tableAligned, tableOff), mr_ptr(scratchAligned, scratchOff), scratchSize, nullptr);
}
int32_t hashJoinWrite(HJ_MEMREF(int32_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch), HJ_MEMREF(int32_t, outR), HJ_MEMREF(int32_t, outS)) {
  (void)SAlloc; (void)tableAlloc; (void)scratchAlloc; (void)outRAlloc; (void)outSAlloc; (void)tableSize; (void)scratchSize;
  const bool ok = MR_ARGS_OK(S) && MR_ARGS_OK(table) && MR_ARGS_OK(scratch) && MR_ARGS_OK(outR) && MR_ARGS_OK(outS);
  return mlir_write(mr_ptr(SAligned, SOff), SSize, ok, 4, mr_ptr(tableAligned, tableOff), mr_ptr(scratchAligned, scratchOff), mr_ptr(outRAligned, outROff), mr_ptr(outSAligned, outSOff));
}
int32_t hashJoinWriteI64(HJ_MEMREF(int64_t, S), HJ_MEMREF(int8_t, table), HJ_MEMREF(int8_t, scratch), HJ_MEMREF(int32_t, outR), HJ_MEMREF(int32_t, outS)) {
  (void)SAlloc; (void)tableAlloc; (void)scratchAlloc; (void)outRAlloc; (void)outSAlloc; (void)tableSize; (void)scratchSize;
  const bool ok = MR_ARGS_OK(S) && MR_ARGS_OK(table) && MR_ARGS_OK(scratch) && MR_ARGS_OK(outR) && MR_ARGS_OK(outS);
  return mlir_write(mr_ptr(SAligned, SOff), SSize, ok, 8, mr_ptr(tableAligned, tableOff), mr_ptr(scratchAligned, scratchOff), mr_ptr(outRAligned, outROff), mr_ptr(outSAligned, outSOff));
}

int32_t hashJoinGather(HJ_MEMREF(int32_t, column), HJ_MEMREF(int32_t, rowIds), HJ_MEMREF(int32_t, out)) {
  (void)columnAlloc; (void)rowIdsAlloc; (void)outAlloc; (void)columnSize;
  if (!(MR_ARGS_OK(column) && MR_ARGS_OK(rowIds) && MR_ARGS_OK(out)) || outSize < rowIdsSize) return fail(HJ_ERR_ARG, "hashJoinGather", "non-unit stride memref or short output");
  int32_t rc = hjGather(mr_ptr(columnAligned, columnOff), 4, mr_ptr(rowIdsAligned, rowIdsOff), rowIdsSize, 0, mr_ptr(outAligned, outOff), nullptr);
  if (rc != HJ_OK) return rc;
  cudaError_t e = cudaStreamSynchronize(nullptr);
  return e == cudaSuccess ? HJ_OK : cuda_fail("hashJoinGather", e);
}

// =========================================================================================================
// B. the reference's join entry points (join_v1.mlir:43-176), expanded ABI.
// The caller's chained-table arrays (head / lkey / lrow / lnext, join_v1.mlir:25-39) are only a HANDLE here: the
// B200 table and its probe-side scratch live in workspaces this library owns, keyed by the head pointer.
// Timer prints are kept where the reference's wrappers have them (join_v1.mlir:65,72,97,105,128,137,164,174).
// =========================================================================================================
int64_t calculateNumberOfBlocks(int64_t totalThreads, int64_t threadsPerBlock) {             // join_v1.mlir:43-52
  return threadsPerBlock > 0 ? (int64_t)(((uint64_t)totalThreads + (uint64_t)threadsPerBlock - 1) / (uint64_t)threadsPerBlock) : 0;
}

void initializeHashTable(int64_t hashTableSize, HJ_MEMREF(int32_t, head)) {                  // join_v1.mlir:54-75
  (void)headAlloc; (void)headStride;
  startTimer();
  int32_t* h = mr_ptr(headAligned, headOff);
  int64_t n = std::min<int64_t>(hashTableSize, headSize);
  if (h && n > 0) {
    cudaError_t e = cudaMemsetAsync(h, 0xFF, (size_t)n * 4, nullptr);                         // head[i] = -1  (:197)
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) cuda_fail("initializeHashTable", e);
  }
  { std::lock_guard<std::mutex> lk(g_mu); g_tables[h].built = false; }
  endTimer();
}

void buildTable(HJ_MEMREF(int32_t, R), int64_t nR, HJ_MEMREF(int32_t, head), HJ_MEMREF(int32_t, lkey), HJ_MEMREF(int64_t, lrow),
                HJ_MEMREF(int64_t, lnext), int32_t hashTableSize) {                          // join_v1.mlir:77-108
  (void)RAlloc; (void)headAlloc; (void)headSize; (void)headStride; (void)lkeyAlloc; (void)lkeyAligned; (void)lkeyOff; (void)lkeySize; (void)lkeyStride;
  (void)lrowAlloc; (void)lrowAligned; (void)lrowOff; (void)lrowSize; (void)lrowStride; (void)lnextAlloc; (void)lnextAligned; (void)lnextOff; (void)lnextSize;
  (void)lnextStride; (void)hashTableSize;
  if (!mr_ok(RSize, RStride) || nR < 0 || nR > RSize) { fail(HJ_ERR_ARG, "buildTable", "bad build relation memref"); return; }
  std::lock_guard<std::mutex> lk(g_mu);
  LegacyTable& t = g_tables[mr_ptr(headAligned, headOff)];
  const int64_t need = hj::table_bytes(nR, 4);
  if (t.table_bytes < need) {
    if (t.table) cudaFree(t.table);
    t.table = nullptr; t.table_bytes = 0;
    cudaError_t e = cudaMalloc(&t.table, (size_t)need);
    if (e != cudaSuccess) { cuda_fail("buildTable", e); return; }
    t.table_bytes = need;
  }
  startTimer();
  int32_t rc = hjBuild(mr_ptr(RAligned, ROff), nR, 4, nullptr, 0, t.table, t.table_bytes, nullptr);
  if (rc == HJ_OK) { cudaError_t e = cudaStreamSynchronize(nullptr); if (e != cudaSuccess) rc = cuda_fail("buildTable", e); }
  t.built = rc == HJ_OK;
  endTimer();
}

int64_t countRows(HJ_MEMREF(int32_t, S), int64_t nS, HJ_MEMREF(int32_t, head), HJ_MEMREF(int32_t, lkey), HJ_MEMREF(int64_t, lrow),
                  HJ_MEMREF(int64_t, lnext), HJ_MEMREF(int64_t, prefix), int32_t hashTableSize) {   // join_v1.mlir:110-147
  (void)SAlloc; (void)headAlloc; (void)headSize; (void)headStride; (void)lkeyAlloc; (void)lkeyAligned; (void)lkeyOff; (void)lkeySize; (void)lkeyStride;
  (void)lrowAlloc; (void)lrowAligned; (void)lrowOff; (void)lrowSize; (void)lrowStride; (void)lnextAlloc; (void)lnextAligned; (void)lnextOff; (void)lnextSize;
  (void)lnextStride; (void)prefixAlloc; (void)hashTableSize;
  if (!mr_ok(SSize, SStride) || nS < 0 || nS > SSize) return fail(HJ_ERR_ARG, "countRows", "bad probe relation memref");
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tables.find(mr_ptr(headAligned, headOff));
  if (it == g_tables.end() || !it->second.built) return fail(HJ_ERR_STATE, "countRows", "hash table was never built (call buildTable first)");
  LegacyTable& t = it->second;
  const int64_t need = hj::scratch_bytes(nS, 4);
  // the caller's prefixSumArray (8 bytes per probe row, join_v1.mlir:588) is never large enough for the probe-side scratch (match cache,
  // offsets, radix area): the library keeps its own, cached with the table
  (void)prefixAligned; (void)prefixOff; (void)prefixSize; (void)prefixStride;
  if (t.scratch_bytes < need) {
    if (t.scratch) cudaFree(t.scratch);
    t.scratch = nullptr; t.scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&t.scratch, (size_t)need);
    if (e != cudaSuccess) return cuda_fail("countRows", e);
    t.scratch_bytes = need;
  }
  startTimer();
  int64_t total = hjCount(mr_ptr(SAligned, SOff), nS, 4, t.table, t.scratch, t.scratch_bytes, nullptr);
  endTimer();
  return total;
}

void probeRelation(HJ_MEMREF(int32_t, S), int64_t nS, int32_t hashTableSize, HJ_MEMREF(int32_t, head), HJ_MEMREF(int32_t, lkey),
                   HJ_MEMREF(int64_t, lrow), HJ_MEMREF(int64_t, lnext), HJ_MEMREF(int64_t, prefix), HJ_MEMREF(int32_t, outR), HJ_MEMREF(int32_t, outS)) {
  (void)SAlloc; (void)hashTableSize; (void)headAlloc; (void)headSize; (void)headStride; (void)lkeyAlloc; (void)lkeyAligned; (void)lkeyOff; (void)lkeySize;
  (void)lkeyStride; (void)lrowAlloc; (void)lrowAligned; (void)lrowOff; (void)lrowSize; (void)lrowStride; (void)lnextAlloc; (void)lnextAligned; (void)lnextOff;
  (void)lnextSize; (void)lnextStride; (void)prefixAlloc; (void)outRAlloc; (void)outSAlloc; (void)outRSize; (void)outSSize;
  if (!mr_ok(SSize, SStride) || !mr_ok(outRSize, outRStride) || !mr_ok(outSSize, outSStride) || nS < 0 || nS > SSize) { fail(HJ_ERR_ARG, "probeRelation", "bad memref"); return; }
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tables.find(mr_ptr(headAligned, headOff));
  if (it == g_tables.end() || !it->second.built) { fail(HJ_ERR_STATE, "probeRelation", "hash table was never built"); return; }
  LegacyTable& t = it->second;
  (void)prefixAligned; (void)prefixOff; (void)prefixSize; (void)prefixStride;
  const void* scratch = t.scratch;
  if (!scratch || t.scratch_bytes < hj::scratch_bytes(nS, 4)) { fail(HJ_ERR_STATE, "probeRelation", "countRows was not called on this table"); return; }
  startTimer();
  int32_t rc = hjWrite(mr_ptr(SAligned, SOff), nS, 4, t.table, scratch, mr_ptr(outRAligned, outROff), mr_ptr(outSAligned, outSOff), nullptr, 0, nullptr);
  if (rc == HJ_OK) { cudaError_t e = cudaStreamSynchronize(nullptr); if (e != cudaSuccess) cuda_fail("probeRelation", e); }
  endTimer();
}

void hashJoinRelease(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_tables) { if (kv.second.table) cudaFree(kv.second.table); if (kv.second.scratch) cudaFree(kv.second.scratch); }
  g_tables.clear();
  g_hR.release(); g_hS.release(); g_hT.release(); g_hSc.release(); for (auto& b : g_hOr) b.release(); for (auto& b : g_hOs) b.release();
}

// =========================================================================================================
// llvm.emit_c_interface wrappers: descriptors by pointer
// =========================================================================================================
#define MR(T, d) (T*)(d)->allocated, (T*)(d)->aligned, (d)->offset, (d)->sizes[0], (d)->strides[0]

void _mlir_ciface_initializeHashTable(int64_t H, HjMemRef1D* head) { initializeHashTable(H, MR(int32_t, head)); }
void _mlir_ciface_buildTable(HjMemRef1D* R, int64_t nR, HjMemRef1D* head, HjMemRef1D* lkey, HjMemRef1D* lrow, HjMemRef1D* lnext, int32_t H) {
  buildTable(MR(int32_t, R), nR, MR(int32_t, head), MR(int32_t, lkey), MR(int64_t, lrow), MR(int64_t, lnext), H);
}
int64_t _mlir_ciface_countRows(HjMemRef1D* S, int64_t nS, HjMemRef1D* head, HjMemRef1D* lkey, HjMemRef1D* lrow, HjMemRef1D* lnext, HjMemRef1D* prefix, int32_t H) {
  return countRows(MR(int32_t, S), nS, MR(int32_t, head), MR(int32_t, lkey), MR(int64_t, lrow), MR(int64_t, lnext), MR(int64_t, prefix), H);
}
void _mlir_ciface_probeRelation(HjMemRef1D* S, int64_t nS, int32_t H, HjMemRef1D* head, HjMemRef1D* lkey, HjMemRef1D* lrow, HjMemRef1D* lnext,
                                HjMemRef1D* prefix, HjMemRef1D* outR, HjMemRef1D* outS) {
  probeRelation(MR(int32_t, S), nS, H, MR(int32_t, head), MR(int32_t, lkey), MR(int64_t, lrow), MR(int64_t, lnext), MR(int64_t, prefix), MR(int32_t, outR), MR(int32_t, outS));
}
int32_t _mlir_ciface_check(HjMemRef1D* R, HjMemRef1D* S, HjMemRef1D* outR, HjMemRef1D* outS) {
  return check(MR(int32_t, R), MR(int32_t, S), MR(int32_t, outR), MR(int32_t, outS));
}
void _mlir_ciface_initRelationIndex(HjMemRef1D* a) { initRelationIndex(MR(int32_t, a)); }
void _mlir_ciface_initRelationR(HjMemRef1D* a) { initRelationR(MR(int32_t, a)); }
void _mlir_ciface_initRelationS(HjMemRef1D* a) { initRelationS(MR(int32_t, a)); }
int32_t _mlir_ciface_hashJoinBuild(HjMemRef1D* R, HjMemRef1D* table) { return hashJoinBuild(MR(int32_t, R), MR(int8_t, table)); }
int64_t _mlir_ciface_hashJoinCount(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch) { return hashJoinCount(MR(int32_t, S), MR(int8_t, table), MR(int8_t, scratch)); }
int32_t _mlir_ciface_hashJoinWrite(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch, HjMemRef1D* outR, HjMemRef1D* outS) {
  return hashJoinWrite(MR(int32_t, S), MR(int8_t, table), MR(int8_t, scratch), MR(int32_t, outR), MR(int32_t, outS));
}
int32_t _mlir_ciface_hashJoinGather(HjMemRef1D* column, HjMemRef1D* rowIds, HjMemRef1D* out) { return hashJoinGather(MR(int32_t, column), MR(int32_t, rowIds), MR(int32_t, out)); }
int32_t _mlir_ciface_hashJoinBuildI64(HjMemRef1D* R, HjMemRef1D* table) { return hashJoinBuildI64(MR(int64_t, R), MR(int8_t, table)); }
int64_t _mlir_ciface_hashJoinCountI64(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch) { return hashJoinCountI64(MR(int64_t, S), MR(int8_t, table), MR(int8_t, scratch)); }
int32_t _mlir_ciface_hashJoinWriteI64(HjMemRef1D* S, HjMemRef1D* table, HjMemRef1D* scratch, HjMemRef1D* outR, HjMemRef1D* outS) {
  return hashJoinWriteI64(MR(int64_t, S), MR(int8_t, table), MR(int8_t, scratch), MR(int32_t, outR), MR(int32_t, outS));
}

}  // extern "C"
