// join_b200_routeB.mlir — MLIR host driver for libhashjoin_b200.so, ROUTE B: the native surface.
//
// Same job as join_v1.mlir:525-658, but the table and the probe-side scratch are two opaque device workspaces
// (memref<?xi8>) sized by query, and the declarations carry `llvm.emit_c_interface`, so each call lowers to
// `_mlir_ciface_<name>(StridedMemRefType<T,1>* ...)` — the descriptor-by-pointer ABI of include/hashjoin_b200.h, group C1.
// The two-phase shape (count -> caller allocates -> write) is the reference's (join_v1.mlir:591,604-605).
// After the join the matched build payloads are materialised with @hashJoinGather (the late gather nested-loop.mlir:165-187
// does inline); negative results are HJ_ERR_* codes.
//
// NOT RUN in the build container (no mlir-opt / mlir-cpu-runner there); the same entry points are driven through the
// descriptor ABI by tests/test_gpu_parity.py::test_native_mlir_surface_ciface.
module attributes {gpu.container_module} {
  memref.global constant @buildRelationRows : memref<1xindex> = dense<[16777216]>
  memref.global constant @probeRelationRows : memref<1xindex> = dense<[268435456]>

  func.func @debugI32(%v : i32) {
    %cell = memref.alloc() : memref<i32>
    memref.store %v, %cell[] : memref<i32>
    %u = memref.cast %cell : memref<i32> to memref<*xi32>
    func.call @printMemrefI32(%u) : (memref<*xi32>) -> ()
    memref.dealloc %cell : memref<i32>
    return
  }

  func.func @main() {
    %c0 = arith.constant 0 : index
    %nRref = memref.get_global @buildRelationRows : memref<1xindex>
    %nSref = memref.get_global @probeRelationRows : memref<1xindex>
    %nR = memref.load %nRref[%c0] : memref<1xindex>
    %nS = memref.load %nSref[%c0] : memref<1xindex>

    %hR = memref.alloc(%nR) : memref<?xi32>
    %hS = memref.alloc(%nS) : memref<?xi32>
    func.call @initRelationR(%hR) : (memref<?xi32>) -> ()
    func.call @initRelationS(%hS) : (memref<?xi32>) -> ()
    %dR = gpu.alloc(%nR) : memref<?xi32>
    %dS = gpu.alloc(%nS) : memref<?xi32>
    gpu.memcpy %dR, %hR : memref<?xi32>, memref<?xi32>
    gpu.memcpy %dS, %hS : memref<?xi32>, memref<?xi32>

    // workspaces: sizes come from the library, memory from the caller (gpu.alloc is 256-byte aligned)
    %tb = func.call @hashJoinTableBytes(%nR) : (index) -> index
    %sb = func.call @hashJoinScratchBytes(%nS) : (index) -> index
    %table = gpu.alloc(%tb) : memref<?xi8>
    %scratch = gpu.alloc(%sb) : memref<?xi8>

    func.call @startTimer() : () -> ()
    %rcB = func.call @hashJoinBuild(%dR, %table) : (memref<?xi32>, memref<?xi8>) -> i32
    %n = func.call @hashJoinCount(%dS, %table, %scratch) : (memref<?xi32>, memref<?xi8>, memref<?xi8>) -> index
    func.call @endTimer() : () -> ()
    %n32 = arith.index_cast %n : index to i32
    func.call @debugI32(%n32) : (i32) -> ()

    %hOutR = memref.alloc(%n) : memref<?xi32>
    %hOutS = memref.alloc(%n) : memref<?xi32>
    %some = arith.cmpi sgt, %n, %c0 : index
    scf.if %some {
      %outR = gpu.alloc(%n) : memref<?xi32>
      %outS = gpu.alloc(%n) : memref<?xi32>
      func.call @startTimer() : () -> ()
      %rcW = func.call @hashJoinWrite(%dS, %table, %scratch, %outR, %outS)
        : (memref<?xi32>, memref<?xi8>, memref<?xi8>, memref<?xi32>, memref<?xi32>) -> i32
      func.call @endTimer() : () -> ()
      // late materialisation: the build-side key of every result row (any i32 payload column works the same way)
      %joined = gpu.alloc(%n) : memref<?xi32>
      %rcG = func.call @hashJoinGather(%dR, %outR, %joined) : (memref<?xi32>, memref<?xi32>, memref<?xi32>) -> i32
      gpu.memcpy %hOutR, %outR : memref<?xi32>, memref<?xi32>
      gpu.memcpy %hOutS, %outS : memref<?xi32>, memref<?xi32>
      gpu.dealloc %joined : memref<?xi32>
      gpu.dealloc %outR : memref<?xi32>
      gpu.dealloc %outS : memref<?xi32>
    }
    %ok = func.call @check(%hR, %hS, %hOutR, %hOutS) : (memref<?xi32>, memref<?xi32>, memref<?xi32>, memref<?xi32>) -> i32
    func.call @debugI32(%ok) : (i32) -> ()
    gpu.dealloc %table : memref<?xi8>
    gpu.dealloc %scratch : memref<?xi8>
    return
  }

  // ---- libhashjoin_b200.so, group A (llvm.emit_c_interface: _mlir_ciface_<name>) ----
  func.func private @initRelationR(memref<?xi32>) attributes { llvm.emit_c_interface }
  func.func private @initRelationS(memref<?xi32>) attributes { llvm.emit_c_interface }
  func.func private @check(memref<?xi32>, memref<?xi32>, memref<?xi32>, memref<?xi32>) -> i32 attributes { llvm.emit_c_interface }
  func.func private @startTimer()
  func.func private @endTimer()
  // ---- group C1: the native surface ----
  func.func private @hashJoinTableBytes(index) -> index
  func.func private @hashJoinScratchBytes(index) -> index
  func.func private @hashJoinBuild(memref<?xi32>, memref<?xi8>) -> i32 attributes { llvm.emit_c_interface }
  func.func private @hashJoinCount(memref<?xi32>, memref<?xi8>, memref<?xi8>) -> index attributes { llvm.emit_c_interface }
  func.func private @hashJoinWrite(memref<?xi32>, memref<?xi8>, memref<?xi8>, memref<?xi32>, memref<?xi32>) -> i32 attributes { llvm.emit_c_interface }
  func.func private @hashJoinGather(memref<?xi32>, memref<?xi32>, memref<?xi32>) -> i32 attributes { llvm.emit_c_interface }
  // ---- libmlir_runner_utils.so ----
  func.func private @printMemrefI32(memref<*xi32>)
}
