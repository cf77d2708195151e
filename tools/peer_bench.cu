// peer_bench.cu — how fast can SMs push a contiguous stream into a PEER GPU's memory over NVLink, by store shape?
//   mode 0: 4 B per lane (128 B per warp store)   mode 1: 8 B per lane (256 B)   mode 2: 16 B per lane (512 B)
//   mode 3: shared memory -> peer by cp.async.bulk (TMA), 4 KB per copy, one thread per CTA issues
// The source is local HBM (mode 0-2: read with the same width; mode 3: staged through shared memory by 16-byte loads).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/libpeer_bench.so tools/peer_bench.cu
#include <cstdint>
#include <cuda_runtime.h>

template <typename V>
__global__ void __launch_bounds__(512) k_push(const V* __restrict__ src, V* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

constexpr int BULK = 4096, STAGES = 4;
__global__ void __launch_bounds__(256) k_push_bulk(const char* __restrict__ src, char* __restrict__ dst, size_t bytes) {
  __shared__ __align__(128) char buf[STAGES][BULK];
  const size_t chunks = bytes / BULK;
  uint32_t it = 0;
  for (size_t c = blockIdx.x; c < chunks; c += gridDim.x, it++) {
    const int s = it % STAGES;
    if (it >= STAGES) {                                   // the copy that last read this stage must have finished reading shared memory
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(STAGES - 1) : "memory");
      __syncthreads();
    }
    reinterpret_cast<int4*>(buf[s])[threadIdx.x] = reinterpret_cast<const int4*>(src + c * BULK)[threadIdx.x];   // 256 threads x 16 B = 4 KB
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(buf[s]);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + c * BULK), "r"(sa), "n"(BULK) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int peer_push(const void* src, void* dst, size_t bytes, int mode, int ctas, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) k_push<uint32_t><<<ctas, 512, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, bytes / 4);
  else if (mode == 1) k_push<uint2><<<ctas, 512, 0, st>>>((const uint2*)src, (uint2*)dst, bytes / 8);
  else if (mode == 2) k_push<int4><<<ctas, 512, 0, st>>>((const int4*)src, (int4*)dst, bytes / 16);
  else k_push_bulk<<<ctas, 256, 0, st>>>((const char*)src, (char*)dst, bytes);
  return (int)cudaGetLastError();
}
