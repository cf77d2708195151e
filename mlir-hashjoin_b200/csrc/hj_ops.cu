// hj_ops.cu — the operators either side of the join path, hand-written for sm_100a. All HBM-bound byte movers.
//
//   late gather / row materialisation   what nested-loop.mlir:165-187 does inline (copy the matched rows of both tables into the result)
//                                        and projectDescription.md:26,30 lists as left out for the hash join ("storing entire data")
//   selection                            Experiments/selection.mlir:34-155: predicate -> count -> scan -> compacted write, same skeleton
//   key packing                          projectDescription.md:28 "multi-columned key join": two i32 columns -> one i64 key (a bijection)
//   column extraction                    the reference's tables are row-major memref<?x?xi32> (nested-loop.mlir:7-24); the join takes columns
#include <algorithm>
#include "hj_common.cuh"
#include "hj_kernels.cuh"

namespace hj {

static inline int64_t r256(int64_t x) { return (x + 255) / 256 * 256; }
static inline unsigned stream_grid(int64_t items, int per_cta) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(148 * 16, (items + per_cta - 1) / per_cta)); }

// ---------------------------------------------------------------------------------------------------------
// late gather: out[i] = column[rows[i] - row_base]      (T = 4- or 8-byte elements; rows = one column of the pair stream)
// The row ids stream (coalesced, evict-first), the column is the random side: 4 loads per thread in flight, the column kept in L2.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BLOCK_THREADS) k_gather(const T* __restrict__ column, const int32_t* __restrict__ rows, int64_t n, uint32_t row_base, T* __restrict__ out) {
  const uint64_t pol = policy_evict_first();
  constexpr int U = 4;
  for (int64_t i0 = (blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x) * U; i0 < n; i0 += (int64_t)gridDim.x * BLOCK_THREADS * U) {
    uint32_t r[U]; T v[U];
    if (i0 + U <= n && (reinterpret_cast<uintptr_t>(rows + i0) & 15) == 0) { const int4 x = ld_stream_v4(rows + i0, pol); r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w; }
    else {
      #pragma unroll
      for (int u = 0; u < U; u++) r[u] = i0 + u < n ? (uint32_t)rows[i0 + u] : row_base;
    }
    #pragma unroll
    for (int u = 0; u < U; u++) v[u] = i0 + u < n ? column[r[u] - row_base] : T(0);
    #pragma unroll
    for (int u = 0; u < U; u++) if (i0 + u < n) out[i0 + u] = v[u];
  }
}
cudaError_t gather_column(const void* column, int elem_bytes, const int32_t* rows, int64_t n, uint32_t row_base, void* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = stream_grid(n, BLOCK_THREADS * 4);
  if (elem_bytes == 4) k_gather<uint32_t><<<grid, BLOCK_THREADS, 0, stream>>>((const uint32_t*)column, rows, n, row_base, (uint32_t*)out);
  else if (elem_bytes == 8) k_gather<unsigned long long><<<grid, BLOCK_THREADS, 0, stream>>>((const unsigned long long*)column, rows, n, row_base, (unsigned long long*)out);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// row materialisation in the reference's layout (nested-loop.mlir:165-187): result row k = all x_cols columns of table x's row
// pair_x[k], then columns 1 .. y_cols-1 of table y's row pair_y[k] (the key is not stored twice). Tables and result are row-major i32.
// One thread per result ELEMENT: the result leaves as full lines, each source row segment is read by neighbouring lanes.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_THREADS) k_materialize_rows(const int32_t* __restrict__ tx, int x_cols, const int32_t* __restrict__ ty, int y_cols,
                                                                    const int32_t* __restrict__ pair_x, const int32_t* __restrict__ pair_y, int64_t n_pairs,
                                                                    int32_t* __restrict__ result) {
  const int out_cols = x_cols + y_cols - 1;
  const int64_t total = n_pairs * out_cols;
  for (int64_t e = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * BLOCK_THREADS) {
    const int64_t k = e / out_cols;
    const int j = (int)(e - k * out_cols);
    result[e] = j < x_cols ? tx[(int64_t)(uint32_t)pair_x[k] * x_cols + j] : ty[(int64_t)(uint32_t)pair_y[k] * y_cols + (j - x_cols + 1)];
  }
}
cudaError_t materialize_rows(const int32_t* tx, int x_cols, const int32_t* ty, int y_cols, const int32_t* pair_x, const int32_t* pair_y, int64_t n_pairs,
                             int32_t* result, cudaStream_t stream) {
  if (x_cols < 1 || y_cols < 1) return cudaErrorInvalidValue;
  if (n_pairs == 0) return cudaSuccess;
  k_materialize_rows<<<stream_grid(n_pairs * (x_cols + y_cols - 1), BLOCK_THREADS * 4), BLOCK_THREADS, 0, stream>>>(tx, x_cols, ty, y_cols, pair_x, pair_y, n_pairs, result);
  return cudaGetLastError();
}

// out[i] = table[i * cols + col]   (row-major memref<?x?xi32> -> key column)
__global__ void __launch_bounds__(BLOCK_THREADS) k_extract_column(const int32_t* __restrict__ table, int64_t rows, int cols, int col, int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; i < rows; i += (int64_t)gridDim.x * BLOCK_THREADS) out[i] = table[i * cols + col];
}
cudaError_t extract_column(const int32_t* table, int64_t rows, int cols, int col, int32_t* out, cudaStream_t stream) {
  if (cols < 1 || col < 0 || col >= cols) return cudaErrorInvalidValue;
  if (rows == 0) return cudaSuccess;
  k_extract_column<<<stream_grid(rows, BLOCK_THREADS), BLOCK_THREADS, 0, stream>>>(table, rows, cols, col, out);
  return cudaGetLastError();
}

// two i32 key columns -> one i64 key: (a, b) == (a', b')  <=>  pack(a, b) == pack(a', b')
__global__ void __launch_bounds__(BLOCK_THREADS) k_pack_keys(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n, long long* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK_THREADS)
    out[i] = (long long)(((unsigned long long)(uint32_t)a[i] << 32) | (uint32_t)b[i]);
}
cudaError_t pack_keys(const int32_t* a, const int32_t* b, int64_t n, long long* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_pack_keys<<<stream_grid(n, BLOCK_THREADS), BLOCK_THREADS, 0, stream>>>(a, b, n, out);
  return cudaGetLastError();
}

// floating-point key columns -> integer keys with IEEE equality: equal numbers get equal keys (-0.0 joins +0.0), a NaN joins nothing
// (build-side NaNs become the positive quiet-NaN pattern, probe-side NaNs the negative one: no number has either, and they differ)
template <typename F, typename I>
__global__ void __launch_bounds__(BLOCK_THREADS) k_encode_float_keys(const F* __restrict__ in, int64_t n, int probe_side, I* __restrict__ out) {
  constexpr I QNAN = sizeof(F) == 4 ? (I)0x7FC00000 : (I)0x7FF8000000000000LL;
  constexpr I SIGN = sizeof(F) == 4 ? (I)0x80000000u : (I)0x8000000000000000ULL;
  for (int64_t i = blockIdx.x * (int64_t)BLOCK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * BLOCK_THREADS) {
    const F v = in[i];
    I bits;
    if (v != v) bits = probe_side ? (I)(QNAN | SIGN) : QNAN;
    else if (v == F(0)) bits = 0;
    else if (sizeof(F) == 4) bits = (I)__float_as_int((float)v);
    else bits = (I)__double_as_longlong((double)v);
    out[i] = bits;
  }
}
cudaError_t encode_float_keys(const void* in, int64_t n, int elem_bytes, int probe_side, void* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (elem_bytes == 4) k_encode_float_keys<float, int32_t><<<stream_grid(n, BLOCK_THREADS), BLOCK_THREADS, 0, stream>>>((const float*)in, n, probe_side, (int32_t*)out);
  else k_encode_float_keys<double, long long><<<stream_grid(n, BLOCK_THREADS), BLOCK_THREADS, 0, stream>>>((const double*)in, n, probe_side, (long long*)out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// selection: rows whose value satisfies `value OP constant`, count -> scan (K3) -> write. A warp owns 512 consecutive rows in both
// passes and compacts them with ballots, so the output keeps the input order (the reference's order depends on which block's
// atomic lands first, selection.mlir:118) and no block barrier is needed.
// ---------------------------------------------------------------------------------------------------------
constexpr int SEL_WARP_ROWS = 512;
template <typename T> __device__ __forceinline__ bool sel_test(T v, int op, T c) {
  switch (op) { case 0: return v < c; case 1: return v <= c; case 2: return v > c; case 3: return v >= c; case 4: return v == c; default: return v != c; }
}
template <typename T>
__global__ void __launch_bounds__(BLOCK_THREADS) k_select_count(const T* __restrict__ col, int64_t n, int op, T c, unsigned long long* __restrict__ totals, int64_t nslices) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (BLOCK_THREADS / 32);
  for (int64_t s = (int64_t)blockIdx.x * (BLOCK_THREADS / 32) + (threadIdx.x >> 5); s < nslices; s += warps) {
    const int64_t base = s * SEL_WARP_ROWS;
    uint32_t cnt = 0;
    #pragma unroll 4
    for (int it = 0; it < SEL_WARP_ROWS / 32; it++) { const int64_t i = base + it * 32 + lane; cnt += i < n && sel_test<T>(col[i], op, c); }
    cnt = warp_reduce_sum(cnt);
    if (lane == 0) totals[s] = cnt;
  }
}
template <typename T>
__global__ void __launch_bounds__(BLOCK_THREADS) k_select_write(const T* __restrict__ col, int64_t n, int op, T c, const unsigned long long* __restrict__ offsets, int64_t nslices,
                                                                T* __restrict__ out_values, int32_t* __restrict__ out_rows, uint32_t row_base) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warps = (int64_t)gridDim.x * (BLOCK_THREADS / 32);
  for (int64_t s = (int64_t)blockIdx.x * (BLOCK_THREADS / 32) + (threadIdx.x >> 5); s < nslices; s += warps) {
    unsigned long long o = offsets[s];
    if (offsets[s + 1] == o) continue;
    const int64_t base = s * SEL_WARP_ROWS;
    #pragma unroll 4
    for (int it = 0; it < SEL_WARP_ROWS / 32; it++) {
      const int64_t i = base + it * 32 + lane;
      const T v = i < n ? col[i] : T(0);
      const bool hit = i < n && sel_test<T>(v, op, c);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const unsigned long long dst = o + __popc(m & lt);
        if (out_values) out_values[dst] = v;
        if (out_rows) out_rows[dst] = (int32_t)(row_base + (uint32_t)i);
      }
      o += __popc(m);
    }
  }
}
// scratch: [offsets u64 x (slices + 1)] [total u64 (+ pad)] [scan block sums u64 x 264]
int64_t select_scratch_bytes(int64_t n) { const int64_t ns = (n + SEL_WARP_ROWS - 1) / SEL_WARP_ROWS; return r256((ns + 1) * 8) + 256 + r256(264 * 8); }
struct SelScratch { unsigned long long* offsets; unsigned long long* total; unsigned long long* sums; int64_t nslices; };
static SelScratch sel_scratch(void* scratch, int64_t n) {
  SelScratch s;
  s.nslices = (n + SEL_WARP_ROWS - 1) / SEL_WARP_ROWS;
  char* p = reinterpret_cast<char*>(scratch);
  s.offsets = reinterpret_cast<unsigned long long*>(p); p += r256((s.nslices + 1) * 8);
  s.total = reinterpret_cast<unsigned long long*>(p); p += 256;
  s.sums = reinterpret_cast<unsigned long long*>(p);
  return s;
}
unsigned long long* select_total_ptr(void* scratch, int64_t n) { return sel_scratch(scratch, n).total; }

template <typename T>
static cudaError_t select_count_t(const T* col, int64_t n, int op, T c, void* scratch, cudaStream_t stream) {
  const SelScratch s = sel_scratch(scratch, n);
  if (s.nslices > 0) k_select_count<T><<<stream_grid(s.nslices, BLOCK_THREADS / 32), BLOCK_THREADS, 0, stream>>>(col, n, op, c, s.offsets, s.nslices);
  launch_scan(s.offsets, s.nslices, s.sums, s.total, stream);
  return cudaGetLastError();
}
template <typename T>
static cudaError_t select_write_t(const T* col, int64_t n, int op, T c, const void* scratch, T* out_values, int32_t* out_rows, uint32_t row_base, cudaStream_t stream) {
  const SelScratch s = sel_scratch(const_cast<void*>(scratch), n);
  if (s.nslices > 0) k_select_write<T><<<stream_grid(s.nslices, BLOCK_THREADS / 32), BLOCK_THREADS, 0, stream>>>(col, n, op, c, s.offsets, s.nslices, out_values, out_rows, row_base);
  return cudaGetLastError();
}
// dtype: 0 = i32, 1 = i64, 2 = f32, 3 = f64; the constant travels as (iconst, fconst)
cudaError_t select_count(const void* col, int64_t n, int dtype, int op, long long iconst, double fconst, void* scratch, cudaStream_t stream) {
  switch (dtype) {
    case 0: return select_count_t<int32_t>((const int32_t*)col, n, op, (int32_t)iconst, scratch, stream);
    case 1: return select_count_t<long long>((const long long*)col, n, op, iconst, scratch, stream);
    case 2: return select_count_t<float>((const float*)col, n, op, (float)fconst, scratch, stream);
    case 3: return select_count_t<double>((const double*)col, n, op, fconst, scratch, stream);
  }
  return cudaErrorInvalidValue;
}
cudaError_t select_write(const void* col, int64_t n, int dtype, int op, long long iconst, double fconst, const void* scratch, void* out_values, int32_t* out_rows, uint32_t row_base,
                         cudaStream_t stream) {
  switch (dtype) {
    case 0: return select_write_t<int32_t>((const int32_t*)col, n, op, (int32_t)iconst, scratch, (int32_t*)out_values, out_rows, row_base, stream);
    case 1: return select_write_t<long long>((const long long*)col, n, op, iconst, scratch, (long long*)out_values, out_rows, row_base, stream);
    case 2: return select_write_t<float>((const float*)col, n, op, (float)fconst, scratch, (float*)out_values, out_rows, row_base, stream);
    case 3: return select_write_t<double>((const double*)col, n, op, fconst, scratch, (double*)out_values, out_rows, row_base, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hj
