// host_driver.cpp — what the reference's lowered @main does (join_v1.mlir:525-649, call ABI witnessed at
// join_v1.ll:983,996,1228,1256,1262-1279), written in C++ because mlir-opt / mlir-cpu-runner are not available here.
// It calls libhashjoin_b200.so exactly the way the JIT-compiled MLIR would: by symbol name, every memref<?xT> expanded
// into (allocated, aligned, offset, size, stride), device buffers from the CUDA runtime (the mgpuMemAlloc wrapper of
// libmlir_cuda_runtime.so is a thin cuMemAlloc).  Usage:
//     hashjoin_main [buildRows probeRows hashTableSize [keyRange]]        (defaults: the 10 x 10, H = 5 snapshot of join_v1.ll:12-14)
// Prints the four timer lines ("For 0..3"), the result size and check()'s verdict like the reference (debugI32 -> printMemrefI32).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hashjoin_b200.h"

#define MEMREF(p, n) (p), (p), (int64_t)0, (int64_t)(n), (int64_t)1

static void debugI32(int32_t v) { printf("Unranked Memref base@ = %p rank = 0 offset = 0 sizes = [] strides = [] data = \n[%d]\n", (void*)&v, v); }   // join_v1.mlir:13-22

template <typename T>
static T* gpu_alloc(int64_t n) {                                            // gpu.alloc -> mgpuMemAlloc
  T* p = nullptr;
  if (cudaMalloc(&p, (size_t)(n > 0 ? n : 1) * sizeof(T)) != cudaSuccess) { fprintf(stderr, "gpu.alloc failed\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  int64_t buildRelationRows = 10, probeRelationRows = 10, hashTableSize = 5;    // join_v1.ll:12-14
  int32_t keyRange = 0;
  if (argc >= 4) { buildRelationRows = atoll(argv[1]); probeRelationRows = atoll(argv[2]); hashTableSize = atoll(argv[3]); }
  if (argc >= 5) keyRange = atoi(argv[4]);

  // :546-549  host relations + generators
  std::vector<int32_t> hostBuild((size_t)buildRelationRows), hostProbe((size_t)probeRelationRows);
  initRelationR(MEMREF(hostBuild.data(), buildRelationRows));
  initRelationS(MEMREF(hostProbe.data(), probeRelationRows));
  if (keyRange > 0) {                                                      // optional: fold keys into [1, keyRange] so that rows match
    for (auto& k : hostBuild) k = k % keyRange + 1;
    for (auto& k : hostProbe) k = k % keyRange + 1;
  }
  // :558-561  device relations
  int32_t* dBuild = gpu_alloc<int32_t>(buildRelationRows);
  int32_t* dProbe = gpu_alloc<int32_t>(probeRelationRows);
  cudaMemcpy(dBuild, hostBuild.data(), (size_t)buildRelationRows * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dProbe, hostProbe.data(), (size_t)probeRelationRows * 4, cudaMemcpyHostToDevice);
  // :565 @allocateHashTable (:25-39): the caller still allocates the reference's four arrays; the library treats them as a handle
  int32_t* lkey = gpu_alloc<int32_t>(buildRelationRows);
  int64_t* lrow = gpu_alloc<int64_t>(buildRelationRows);
  int64_t* lnext = gpu_alloc<int64_t>(buildRelationRows);
  int32_t* head = gpu_alloc<int32_t>(hashTableSize);
  // :568-571
  initializeHashTable(hashTableSize, MEMREF(head, hashTableSize));
  buildTable(MEMREF(dBuild, buildRelationRows), buildRelationRows, MEMREF(head, hashTableSize), MEMREF(lkey, buildRelationRows),
             MEMREF(lrow, buildRelationRows), MEMREF(lnext, buildRelationRows), (int32_t)hashTableSize);
  // :588-597
  int64_t* prefixSumArray = gpu_alloc<int64_t>(probeRelationRows);
  int64_t resultSize = countRows(MEMREF(dProbe, probeRelationRows), probeRelationRows, MEMREF(head, hashTableSize), MEMREF(lkey, buildRelationRows),
                                 MEMREF(lrow, buildRelationRows), MEMREF(lnext, buildRelationRows), MEMREF(prefixSumArray, probeRelationRows),
                                 (int32_t)hashTableSize);
  if (resultSize < 0) { fprintf(stderr, "countRows failed: %s\n", hjLastErrorString()); return 3; }
  debugI32((int32_t)resultSize);
  int32_t success;
  if (resultSize != 0) {                                                   // :600-632
    int32_t* dOutR = gpu_alloc<int32_t>(resultSize);
    int32_t* dOutS = gpu_alloc<int32_t>(resultSize);
    probeRelation(MEMREF(dProbe, probeRelationRows), probeRelationRows, (int32_t)hashTableSize, MEMREF(head, hashTableSize),
                  MEMREF(lkey, buildRelationRows), MEMREF(lrow, buildRelationRows), MEMREF(lnext, buildRelationRows),
                  MEMREF(prefixSumArray, probeRelationRows), MEMREF(dOutR, resultSize), MEMREF(dOutS, resultSize));
    std::vector<int32_t> hostR((size_t)resultSize), hostS((size_t)resultSize);
    cudaMemcpy(hostR.data(), dOutR, (size_t)resultSize * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hostS.data(), dOutS, (size_t)resultSize * 4, cudaMemcpyDeviceToHost);
    success = check(MEMREF(hostBuild.data(), buildRelationRows), MEMREF(hostProbe.data(), probeRelationRows),
                    MEMREF(hostR.data(), resultSize), MEMREF(hostS.data(), resultSize));
  } else {                                                                 // :635-644
    int32_t dummy = 0;
    success = check(MEMREF(hostBuild.data(), buildRelationRows), MEMREF(hostProbe.data(), probeRelationRows), MEMREF(&dummy, 0), MEMREF(&dummy, 0));
  }
  debugI32(success);
  hashJoinRelease();
  return success == 1 ? 0 : 1;
}
