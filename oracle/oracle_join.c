/*
 * oracle_join.c — CPU ORACLE for the hash-join hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path (the CUDA library
 * under mlir-hashjoin_b200/csrc) never links, imports or falls back to anything in oracle/.
 *
 * It restates, as plain C loops, the algorithm of deveshv-99/mlir-HashJoin for the path
 * BASELINE.json names (citations are file:line under /root/reference):
 *
 *   oracle_check_*      <- shared_stuff/shared.cpp:129-172   nested-loop join + sort + compare (THE definition
 *                                                            of a correct result; tri-state 1 / 0 / -1)
 *   oracle_v1_init      <- join_v1.mlir:180-202              head[i] = -1
 *   oracle_v1_hash      <- join_v1.mlir:206-210              (uint32)key % H            (arith.remui)
 *   oracle_v1_build     <- join_v1.mlir:213-249, 251-277     node = free++; key,row; old = xchg(head[h], node); next = old
 *   oracle_v1_count     <- join_v1.mlir:288-425              chain walk, per-row count, exclusive scan -> prefix, total
 *   oracle_v1_probe     <- join_v1.mlir:436-521              second chain walk, outR[w]=(i32)row, outS[w]=tid, w++
 *   oracle_v1_join      <- join_v1.mlir:525-649              the @main sequence init -> build -> count -> probe
 *   oracle_nested_join  <- nested-loop.mlir:78-188           O(n*m) compare, count -> scan -> write skeleton
 *   oracle_nested_join_rows <- nested-loop.mlir:165-187      the same loops, materialising the joined rows
 *   oracle_select_*     <- Experiments/selection.mlir:34-155 predicate -> count -> scan -> compacted write
 *   oracle_init_index   <- shared_stuff/shared.cpp:35-41     a[i] = i
 *
 * Parity pinning: the reference's own check() compiles here (oracle/Makefile -> oracle/_ref/shared.so) and
 * tests/golden/ holds verdicts it produced for the reference's derivable fixtures;
 * tests/test_oracle.py replays them against this file, and against _ref directly when it is present.
 * int64-key variants are type-widened restatements: the reference is i32-only (join_v1.mlir:77), so for
 * int64 parity is pinned only through this chain (stated in DESIGN.md).
 *
 * The seeded generators at the bottom are NOT reference code (the reference's are unseeded rand(),
 * shared.cpp:62,86-87); they exist so CPU and GPU see bit-identical inputs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * check(): shared.cpp:129-172
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t r, s; } pair32;

static int pair_cmp(const void* a, const void* b) {
  const pair32* x = (const pair32*)a; const pair32* y = (const pair32*)b;   /* std::pair operator< : first, then second */
  if (x->r != y->r) return x->r < y->r ? -1 : 1;
  if (x->s != y->s) return x->s < y->s ? -1 : 1;
  return 0;
}

#define DEFINE_CHECK(NAME, KEY_T)                                                                              \
  API int32_t NAME(const KEY_T* R, int64_t nR, const KEY_T* S, int64_t nS,                                     \
                   const int32_t* outR, const int32_t* outS, int64_t result_size) {                            \
    /* shared.cpp:140-141: both vectors are value-initialised to (0,0) with result_size elements */            \
    pair32* got = (pair32*)calloc((size_t)(result_size > 0 ? result_size : 1), sizeof(pair32));                \
    pair32* ref = (pair32*)calloc((size_t)(result_size > 0 ? result_size : 1), sizeof(pair32));                \
    for (int64_t i = 0; i < result_size; i++) { got[i].r = outR[i]; got[i].s = outS[i]; } /* :147-149 */       \
    int64_t cur = 0;                                                                                           \
    for (int64_t i = 0; i < nR; ++i) {                                           /* :154 */                    \
      for (int64_t j = 0; j < nS; ++j) {                                         /* :155 */                    \
        if (R[i] == S[j]) {                                                      /* :156 */                    \
          if (cur >= result_size) { free(got); free(ref); return -1; }           /* :158-160 */                \
          ref[cur].r = (int32_t)i; ref[cur].s = (int32_t)j; cur++;               /* :161-162 */                \
        }                                                                                                      \
      }                                                                                                        \
    }                                                                                                          \
    qsort(ref, (size_t)result_size, sizeof(pair32), pair_cmp);                   /* :168 */                    \
    qsort(got, (size_t)result_size, sizeof(pair32), pair_cmp);                   /* :169 */                    \
    int32_t eq = memcmp(ref, got, (size_t)result_size * sizeof(pair32)) == 0;    /* :171 */                    \
    free(got); free(ref);                                                                                      \
    return eq;                                                                                                 \
  }
DEFINE_CHECK(oracle_check_i32, int32_t)
DEFINE_CHECK(oracle_check_i64, int64_t)

/* Same 20-scalar expanded-memref ABI as the reference export (shared.cpp:129-132), so that tests can call
 * this and oracle/_ref/shared.so:check interchangeably. */
API int32_t oracle_check_memref(int32_t* rBase, int32_t* rAligned, int64_t rOff, int64_t rSize, int64_t rStride,
                                int32_t* sBase, int32_t* sAligned, int64_t sOff, int64_t sSize, int64_t sStride,
                                int32_t* orBase, int32_t* orAligned, int64_t orOff, int64_t orSize, int64_t orStride,
                                int32_t* osBase, int32_t* osAligned, int64_t osOff, int64_t osSize, int64_t osStride) {
  (void)rBase; (void)rOff; (void)rStride; (void)sBase; (void)sOff; (void)sStride;
  (void)orBase; (void)orOff; (void)orStride; (void)osBase; (void)osOff; (void)osSize; (void)osStride;
  return oracle_check_i32(rAligned, rSize, sAligned, sSize, orAligned, osAligned, orSize);
}

/* ------------------------------------------------------------------------------------------------
 * join_v1 as loops.  Table layout is the reference's (join_v1.mlir:25-39):
 *   head  i32[H], lkey KEY[nR], lrow i64[nR], lnext i64[nR]
 * threads > 1 maps memref.atomic_rmw to __atomic builtins (node order then depends on scheduling,
 * exactly as on the GPU, join_v1.mlir:224); threads == 1 is the deterministic serial order.
 * ---------------------------------------------------------------------------------------------- */
API void oracle_v1_init(int32_t* head, int64_t H) {                 /* join_v1.mlir:180-202 */
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < H; i++) head[i] = -1;
}

static inline uint32_t v1_hash32(int32_t key, uint32_t H) { return (uint32_t)key % H; }          /* :206-210 */
static inline uint32_t v1_hash64(int64_t key, uint32_t H) { return (uint32_t)((uint64_t)key % H); } /* widened */

#define DEFINE_V1(SUF, KEY_T, HASH)                                                                            \
  API void oracle_v1_build_##SUF(const KEY_T* R, int64_t nR, int32_t* head, KEY_T* lkey, int64_t* lrow,        \
                                 int64_t* lnext, int32_t H) {                                                  \
    int32_t free_index = 0;                                            /* join_v1.mlir:86-89 */                \
    _Pragma("omp parallel for schedule(static)")                                                               \
    for (int64_t tid = 0; tid < nR; tid++) {                           /* :261-265 one thread per build row */ \
      KEY_T key = R[tid];                                              /* :269 */                              \
      int32_t n = __atomic_fetch_add(&free_index, 1, __ATOMIC_RELAXED);/* :224 */                              \
      lkey[n] = key;                                                   /* :231 */                              \
      lrow[n] = tid;                                                   /* :232 */                              \
      uint32_t h = HASH(key, (uint32_t)H);                             /* :235 */                              \
      int32_t old = __atomic_exchange_n(&head[h], n, __ATOMIC_RELAXED);/* :243 */                              \
      lnext[n] = (int64_t)old;                                         /* :246 */                              \
    }                                                                                                          \
  }                                                                                                            \
  API int64_t oracle_v1_count_##SUF(const KEY_T* S, int64_t nS, const int32_t* head, const KEY_T* lkey,        \
                                    const int64_t* lnext, int64_t* prefix, int32_t H) {                        \
    _Pragma("omp parallel for schedule(static)")                                                               \
    for (int64_t tid = 0; tid < nS; tid++) {                                                                   \
      KEY_T key = S[tid];                                              /* :324 */                              \
      int64_t cnt = 0;                                                 /* :316-319 */                          \
      int64_t cur = head[HASH(key, (uint32_t)H)];                      /* :327-330 */                          \
      while (cur != -1) {                                              /* :333-362 */                          \
        if (lkey[cur] == key) cnt++;                                   /* :344-353 */                          \
        cur = lnext[cur];                                              /* :357 */                              \
      }                                                                                                        \
      prefix[tid] = cnt;                                                                                       \
    }                                                                                                          \
    /* :371-420: block-local exclusive scan + global block offset.  Block bases are handed out in atomic      \
     * arrival order on the GPU; any order yields disjoint ranges, the oracle uses row order. */               \
    int64_t run = 0;                                                                                           \
    for (int64_t tid = 0; tid < nS; tid++) { int64_t c = prefix[tid]; prefix[tid] = run; run += c; }           \
    return run;                                                        /* :140-146 result size */              \
  }                                                                                                            \
  API void oracle_v1_probe_##SUF(const KEY_T* S, int64_t nS, const int32_t* head, const KEY_T* lkey,           \
                                 const int64_t* lrow, const int64_t* lnext, const int64_t* prefix,             \
                                 int32_t* outR, int32_t* outS, int32_t H) {                                    \
    _Pragma("omp parallel for schedule(static)")                                                               \
    for (int64_t tid = 0; tid < nS; tid++) {                                                                   \
      KEY_T key = S[tid];                                              /* :469 */                              \
      int64_t w = prefix[tid];                                         /* :475-476 */                          \
      int64_t cur = head[HASH(key, (uint32_t)H)];                                                              \
      while (cur != -1) {                                              /* :483-514 */                          \
        if (lkey[cur] == key) {                                                                                \
          outR[w] = (int32_t)lrow[cur];                                /* :492-498 */                          \
          outS[w] = (int32_t)tid;                                      /* :499-500 */                          \
          w++;                                                         /* :503 */                              \
        }                                                                                                      \
        cur = lnext[cur];                                                                                      \
      }                                                                                                        \
    }                                                                                                          \
  }                                                                                                            \
  /* @main sequence (join_v1.mlir:525-649) with caller-owned result buffers.  Two-phase like the reference:    \
   * call with outR == NULL to get the result size (countRows, :591), allocate, call again to fill. */         \
  API int64_t oracle_v1_join_##SUF(const KEY_T* R, int64_t nR, const KEY_T* S, int64_t nS, int32_t H,          \
                                   int32_t* outR, int32_t* outS, int64_t capacity, int threads) {              \
    if (H <= 0) return -2;                                                                                     \
    int prev_threads = 1; (void)prev_threads;                                                                  \
    OMP_SET(threads)                                                                                           \
    int32_t* head = (int32_t*)malloc(sizeof(int32_t) * (size_t)H);                                             \
    KEY_T* lkey = (KEY_T*)malloc(sizeof(KEY_T) * (size_t)(nR > 0 ? nR : 1));                                   \
    int64_t* lrow = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nR > 0 ? nR : 1));                             \
    int64_t* lnext = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nR > 0 ? nR : 1));                            \
    int64_t* prefix = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nS > 0 ? nS : 1));                           \
    int64_t total = -3;                                                                                        \
    if (head && lkey && lrow && lnext && prefix) {                                                             \
      oracle_v1_init(head, H);                                                                                 \
      oracle_v1_build_##SUF(R, nR, head, lkey, lrow, lnext, H);                                                \
      total = oracle_v1_count_##SUF(S, nS, head, lkey, lnext, prefix, H);                                      \
      if (outR && outS && total != 0 && total <= capacity)             /* :600-601 zero-size skips probe */    \
        oracle_v1_probe_##SUF(S, nS, head, lkey, lrow, lnext, prefix, outR, outS, H);                          \
    }                                                                                                          \
    free(head); free(lkey); free(lrow); free(lnext); free(prefix);                                             \
    OMP_RESTORE()                                                                                              \
    return total;                                                                                              \
  }

#ifdef _OPENMP
#define OMP_SET(t) prev_threads = omp_get_max_threads(); if ((t) > 0) omp_set_num_threads(t);
#define OMP_RESTORE() omp_set_num_threads(prev_threads);
#else
#define OMP_SET(t) (void)(t);
#define OMP_RESTORE()
#endif

DEFINE_V1(i32, int32_t, v1_hash32)
DEFINE_V1(i64, int64_t, v1_hash64)

API int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* nested-loop.mlir:78-188 — count -> scan -> write with a row-id pair result (the hash join inherited this
 * skeleton).  Output order: probe-major is NOT what the reference does; it is thread = row of table 1
 * (nested-loop.mlir:95-110), inner loop over table 2, so pairs come out build-major. */
API int64_t oracle_nested_join_i32(const int32_t* R, int64_t nR, const int32_t* S, int64_t nS,
                                   int32_t* outR, int32_t* outS, int64_t capacity) {
  int64_t total = 0;
  for (int64_t i = 0; i < nR; i++)
    for (int64_t j = 0; j < nS; j++)
      if (R[i] == S[j]) {
        if (outR && outS && total < capacity) { outR[total] = (int32_t)i; outS[total] = (int32_t)j; }
        total++;
      }
  return total;
}

/* nested-loop.mlir:165-187 — the same loops with ROW MATERIALISATION: result row = all x_cols columns of table x's row i, then
 * columns 1 .. y_cols-1 of table y's row j (the key, column 0, is not stored twice: the inner loop starts at j = 1, :181).
 * Tables and result are row-major i32 (memref<?x?xi32>, :7-24). Returns the number of result rows; fills at most `capacity`. */
API int64_t oracle_nested_join_rows_i32(const int32_t* tx, int64_t x_rows, int32_t x_cols, const int32_t* ty, int64_t y_rows, int32_t y_cols,
                                        int32_t* result, int64_t capacity) {
  const int64_t out_cols = (int64_t)x_cols + y_cols - 1;
  int64_t total = 0;
  for (int64_t i = 0; i < x_rows; i++)                                           /* thread = row of table x, :95-110 */
    for (int64_t j = 0; j < y_rows; j++)                                         /* :160 */
      if (tx[i * x_cols] == ty[j * y_cols]) {                                    /* :161-163: keys are column 0 */
        if (result && total < capacity) {
          for (int32_t c = 0; c < x_cols; c++) result[total * out_cols + c] = tx[i * x_cols + c];                    /* :170-173 */
          for (int32_t c = 1; c < y_cols; c++) result[total * out_cols + x_cols - 1 + c] = ty[j * y_cols + c];       /* :179-184 */
        }
        total++;
      }
  return total;
}

/* Experiments/selection.mlir:34-155 — predicate, count, scan, compacted write (f32 `olt` there: ordered less-than, :62; the
 * other comparisons and types are the same loop). op: 0 <, 1 <=, 2 >, 3 >=, 4 ==, 5 !=. Output in input order (what one block of
 * the reference produces, :140-152; across blocks the reference's order depends on which atomic lands first, :118). */
#define DEFINE_SELECT(NAME, T)                                                                                   \
  API int64_t NAME(const T* col, int64_t n, int32_t op, T c, T* out_values, int32_t* out_rows, int64_t capacity) { \
    int64_t total = 0;                                                                                           \
    for (int64_t i = 0; i < n; i++) {                                                                            \
      const T v = col[i];                                                                                        \
      const int keep = op == 0 ? v < c : op == 1 ? v <= c : op == 2 ? v > c : op == 3 ? v >= c : op == 4 ? v == c : v != c; \
      if (keep) {                                                                                                \
        if (total < capacity) { if (out_values) out_values[total] = v; if (out_rows) out_rows[total] = (int32_t)i; } \
        total++;                                                                                                 \
      }                                                                                                          \
    }                                                                                                            \
    return total;                                                                                                \
  }
DEFINE_SELECT(oracle_select_i32, int32_t)
DEFINE_SELECT(oracle_select_i64, int64_t)
DEFINE_SELECT(oracle_select_f32, float)
DEFINE_SELECT(oracle_select_f64, double)

API void oracle_init_index(int32_t* a, int64_t n) { for (int64_t i = 0; i < n; i++) a[i] = (int32_t)i; } /* shared.cpp:35-41 */

/* ------------------------------------------------------------------------------------------------
 * Order-independent digest of a pair stream (count is returned by the join; this adds sum and xor of a
 * 64-bit mix of every pair) — used for full-size parity where sorting 2^28 pairs on the host is too slow.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {            /* splitmix64 finaliser */
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z;
}
API void oracle_pair_digest(const int32_t* outR, const int32_t* outS, int64_t n, uint64_t* sum_out, uint64_t* xor_out) {
  uint64_t s = 0, x = 0;
  #pragma omp parallel for schedule(static) reduction(+ : s) reduction(^ : x)
  for (int64_t i = 0; i < n; i++) {
    uint64_t m = mix64(((uint64_t)(uint32_t)outR[i] << 32) | (uint32_t)outS[i]);
    s += m; x ^= m;
  }
  *sum_out = s; *xor_out = x;
}

/* ------------------------------------------------------------------------------------------------
 * Seeded generators (integer-only so that the CUDA generators in csrc/datagen.cu are bit-identical).
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t mix32(uint32_t x) {            /* lowbias32 */
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
static inline uint64_t rnd64(uint64_t seed, uint64_t i) { return mix64(seed * 0x9E3779B97F4A7C15ULL + mix64(i + 0xD1B54A32D192ED03ULL)); }

/* Feistel bijection on [0, n): 4 rounds over 2*hb bits, cycle-walking back into range. */
static inline int half_bits(uint64_t n) { int b = 1; while (b < 64 && ((uint64_t)1 << b) < n) b++; return (b + 1) / 2; }
static inline uint64_t feistel_once(uint64_t x, int hb, uint64_t seed, int inverse) {
  uint64_t mask = ((uint64_t)1 << hb) - 1;
  uint64_t l = x >> hb, r = x & mask;
  for (int k = 0; k < 4; k++) {
    int rk = inverse ? 3 - k : k;
    uint64_t f = mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(rk + 1));
    if (!inverse) { uint64_t t = l ^ (mix64(r ^ f) & mask); l = r; r = t; }
    else          { uint64_t t = r ^ (mix64(l ^ f) & mask); r = l; l = t; }
  }
  return (l << hb) | r;
}
API uint64_t oracle_perm(uint64_t i, uint64_t n, uint64_t seed) {
  int hb = half_bits(n); uint64_t x = i;
  do { x = feistel_once(x, hb, seed, 0); } while (x >= n);
  return x;
}
API uint64_t oracle_perm_inv(uint64_t y, uint64_t n, uint64_t seed) {
  int hb = half_bits(n); uint64_t x = y;
  do { x = feistel_once(x, hb, seed, 1); } while (x >= n);
  return x;
}

/* Zipf(1.0) rank in [1, D) by inverse-CDF of the 1/x density, integer only: octave uniform, then 2^f from a
 * 257-entry table built by integer multiplication (no libm, so CPU and GPU agree bit for bit). */
#define ZIPF_C 0x8058D7D2D5E5F6B1ULL    /* round(2^(1/256) * 2^63) */
API void oracle_zipf_table(uint64_t* t) {   /* t[i] = 2^(i/256) in 2.62 fixed point */
  t[0] = (uint64_t)1 << 62;
  for (int i = 1; i <= 256; i++) t[i] = (uint64_t)(((__uint128_t)t[i - 1] * ZIPF_C) >> 63);
  t[256] = (uint64_t)1 << 63;
}
static inline uint64_t zipf_rank(uint64_t r, int log2D, const uint64_t* t) {
  uint32_t a = (uint32_t)(((r >> 32) * (uint64_t)log2D) >> 32);          /* octave 0 .. log2D-1 */
  uint32_t f = (uint32_t)r & 0xFFFF;                                     /* 16-bit fraction */
  uint32_t hi = f >> 8, lo = f & 0xFF;
  uint64_t m = t[hi] + (((t[hi + 1] - t[hi]) * lo) >> 8);                /* 2^f, 2.62 fixed point */
  return (m >> (62 - a));                                                /* floor(2^(a+f)) in [2^a, 2^(a+1)) */
}

/* kind: 0 index (a[i]=i, shared.cpp:35-41)
 *       1 unique:   lo + perm(i, n_domain)            (needs n <= domain)
 *       2 uniform:  lo + floor(u * domain)
 *       3 mixed:    with probability p16/65536 a key from [lo, lo+domain) else from [lo+domain, lo+2*domain)
 *       4 fk:       lo + (perm(i, n) mod domain)      (each key n/domain times when domain | n, rows shuffled)
 *       5 zipf:     lo + perm(zipf_rank - 1, domain)  (domain must be a power of two)
 * key_mul != 0 spreads keys over the full key space by an odd multiplier (bijective mod 2^32 / 2^64). */
static inline int64_t gen_one(int kind, uint64_t i, uint64_t n, uint64_t seed, int64_t lo, uint64_t domain, uint32_t p16,
                              const uint64_t* zt, int log2D) {
  uint64_t v;
  switch (kind) {
    case 0: v = i; break;
    case 1: v = oracle_perm(i, domain, seed); break;
    case 2: v = (uint64_t)(((__uint128_t)rnd64(seed, i) * domain) >> 64); break;
    case 3: { uint64_t r = rnd64(seed, i); uint64_t u = (uint64_t)(((__uint128_t)rnd64(seed ^ 0xA5A5A5A5ULL, i) * domain) >> 64);
              v = ((r & 0xFFFF) < p16) ? u : domain + u; break; }
    case 4: v = oracle_perm(i, n, seed) % domain; break;
    case 5: v = oracle_perm(zipf_rank(rnd64(seed, i), log2D, zt) - 1, domain, seed ^ 0x5EEDULL); break;
    default: v = 0;
  }
  return lo + (int64_t)v;
}
static int ilog2(uint64_t d) { int l = 0; while (((uint64_t)1 << (l + 1)) <= d) l++; return l; }

API void oracle_gen_i32(int32_t* out, int64_t n, int kind, uint64_t seed, int64_t lo, uint64_t domain, uint32_t p16, uint64_t key_mul) {
  uint64_t zt[257]; oracle_zipf_table(zt); int l2 = ilog2(domain ? domain : 1);
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t v = gen_one(kind, (uint64_t)i, (uint64_t)n, seed, lo, domain, p16, zt, l2);
    out[i] = key_mul ? (int32_t)((uint32_t)v * (uint32_t)key_mul) : (int32_t)v;
  }
}
API void oracle_gen_i64(int64_t* out, int64_t n, int kind, uint64_t seed, int64_t lo, uint64_t domain, uint32_t p16, uint64_t key_mul) {
  uint64_t zt[257]; oracle_zipf_table(zt); int l2 = ilog2(domain ? domain : 1);
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t v = gen_one(kind, (uint64_t)i, (uint64_t)n, seed, lo, domain, p16, zt, l2);
    out[i] = key_mul ? (int64_t)((uint64_t)v * key_mul) : v;
  }
}
