#!/bin/bash
# Lower and run one of the join_b200_route*.mlir drivers against libhashjoin_b200.so.
# Same tool chain as the reference's run_test.sh (mlir-opt | mlir-cpu-runner, LLVM 16-17 era pass names), minus the device-code
# passes (gpu-kernel-outlining, convert-gpu-to-nvvm, gpu-to-cubin): the drivers contain no gpu.func, the kernels are in the library.
# Needs an LLVM/MLIR build with the CUDA runner (not available in this repo's build container).
set -euo pipefail
HERE=$(dirname "$(realpath -s "$0")")
: "${LLVM_BUILD_DIR:=$HOME/llvm-project/build}"
LIB=$HERE/../mlir-hashjoin_b200/lib/libhashjoin_b200.so
[ -f "$LIB" ] || { echo "build the library first: python -c 'import __graft_entry__ as g; g.build()'" >&2; exit 1; }
OPT=$LLVM_BUILD_DIR/bin/mlir-opt
"$OPT" -convert-scf-to-cf "${1:-$HERE/join_b200_routeA.mlir}" \
  | "$OPT" -gpu-async-region -arith-expand -convert-arith-to-llvm -convert-cf-to-llvm -finalize-memref-to-llvm -convert-func-to-llvm -gpu-to-llvm -reconcile-unrealized-casts \
  | "$LLVM_BUILD_DIR/bin/mlir-cpu-runner" --shared-libs="$LLVM_BUILD_DIR/lib/libmlir_cuda_runtime.so" --shared-libs="$LLVM_BUILD_DIR/lib/libmlir_runner_utils.so" \
      --shared-libs="$LLVM_BUILD_DIR/lib/libmlir_async_runtime.so" --shared-libs="$LIB" --entry-point-result=void -O0
