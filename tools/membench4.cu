// membench4.cu — what does one table insert cost?  2^28 random CAS.128 / CAS.64 / plain 16-byte stores / 32-byte loads into a table
// of 32 MB (L2-resident) or 1 GB, four independent operations per thread in flight.  (the i64 build is bound by
// the L2's 128-bit compare-and-swap rate, not by DRAM.)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membench4 tools/membench4.cu
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t z) { z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31; return z; }
__device__ __forceinline__ ulonglong2 cas128(ulonglong2* addr, ulonglong2 cmp, ulonglong2 val) {
  ulonglong2 old;
  asm volatile("{\n .reg .b128 c, v, o;\n mov.b128 c, {%2, %3};\n mov.b128 v, {%4, %5};\n atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n mov.b128 {%0, %1}, o;\n}"
               : "=l"(old.x), "=l"(old.y) : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr) : "memory");
  return old;
}
// MODE 0: CAS.128  1: CAS.64  2: st.v2.u64 (16 B)  3: ld.v4.u64 (32 B)  4: ld 32 B then CAS.128 on the same bucket (the insert as built today)
template <int MODE>
__global__ void __launch_bounds__(256) k_ops(char* table, uint64_t slots16_mask, size_t n, unsigned long long* sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint64_t h[4];
    #pragma unroll
    for (int u = 0; u < 4; u++) h[u] = mix64(i + u * stride + 12345);
    if (MODE == 0) {
      ulonglong2 o[4];
      #pragma unroll
      for (int u = 0; u < 4; u++) o[u] = cas128(reinterpret_cast<ulonglong2*>(table) + (h[u] & slots16_mask), make_ulonglong2(~0ULL, ~0ULL), make_ulonglong2(h[u], i));
      #pragma unroll
      for (int u = 0; u < 4; u++) acc += o[u].x;
    } else if (MODE == 1) {
      unsigned long long o[4];
      #pragma unroll
      for (int u = 0; u < 4; u++) o[u] = atomicCAS(reinterpret_cast<unsigned long long*>(table) + 2 * (h[u] & slots16_mask), ~0ULL, h[u]);
      #pragma unroll
      for (int u = 0; u < 4; u++) acc += o[u];
    } else if (MODE == 2) {
      #pragma unroll
      for (int u = 0; u < 4; u++) reinterpret_cast<ulonglong2*>(table)[h[u] & slots16_mask] = make_ulonglong2(h[u], i);
    } else if (MODE == 3) {
      #pragma unroll
      for (int u = 0; u < 4; u++) {
        unsigned long long a, b, c, d;
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(table + 32 * ((h[u] & slots16_mask) >> 1)));
        acc += a ^ b ^ c ^ d;
      }
    } else {
      unsigned long long w[4][4];
      #pragma unroll
      for (int u = 0; u < 4; u++) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[u][0]), "=l"(w[u][1]), "=l"(w[u][2]), "=l"(w[u][3]) : "l"(table + 32 * ((h[u] & slots16_mask) >> 1)));
      ulonglong2 o[4];
      #pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = (uint32_t)w[u][1] == 0xFFFFFFFFu ? 0 : 1;
        o[u] = cas128(reinterpret_cast<ulonglong2*>(table + 32 * ((h[u] & slots16_mask) >> 1)) + e, make_ulonglong2(~0ULL, ~0ULL), make_ulonglong2(h[u], i));
      }
      #pragma unroll
      for (int u = 0; u < 4; u++) acc += o[u].x;
    }
  }
  if (acc == 0x1234567ULL) *sink = acc;
}
template <int MODE>
static void run(const char* name, char* table, size_t table_bytes, size_t n, unsigned long long* sink, int per_sm) {
  float best = 1e30f;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int r = 0; r < 3; r++) {
    CK(cudaMemset(table, 0xFF, table_bytes));
    CK(cudaEventRecord(a));
    k_ops<MODE><<<148 * per_sm, 256>>>(table, table_bytes / 16 - 1, n, sink);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  printf("{\"bench\": \"%s\", \"table_MB\": %zu, \"ops\": %zu, \"ctas_per_sm\": %d, \"ms\": %.3f, \"G_ops_per_s\": %.1f}\n", name, table_bytes >> 20, n, per_sm, best, n / best / 1e6);
  fflush(stdout);
}
int main() {
  const size_t n = (size_t)1 << 28;
  char* table; unsigned long long* sink;
  CK(cudaMalloc(&table, (size_t)8 << 30)); CK(cudaMalloc(&sink, 8));
  for (size_t mb : {32, 1024, 8192}) {
    const size_t bytes = mb << 20;
    run<0>("cas128", table, bytes, n, sink, 8);
    run<1>("cas64", table, bytes, n, sink, 8);
    run<2>("st128", table, bytes, n, sink, 8);
    run<3>("ld256", table, bytes, n, sink, 8);
    run<4>("ld256_then_cas128", table, bytes, n, sink, 8);
  }
  run<0>("cas128", table, (size_t)32 << 20, n, sink, 4);
  run<0>("cas128", table, (size_t)32 << 20, n, sink, 2);
  return 0;
}
