"""In-tree build of libhashjoin_b200.so (nvcc, sm_100a only) and of the C++ host driver."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libhashjoin_b200.so"
DRIVER = PKG / "lib" / "hashjoin_main"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def library_path() -> Path:
    return LIB


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libhashjoin_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into lib/libhashjoin_b200.so for sm_100a. Returns the path."""
    srcs = sorted(CSRC.glob("hj_*.cu"))
    deps = srcs + [CSRC / "hj_common.cuh", CSRC / "hj_kernels.cuh", PKG.parent / "include" / "hashjoin_b200.h"]
    LIB.parent.mkdir(parents=True, exist_ok=True)
    if force or _stale(LIB, deps):
        # one nvcc per translation unit, side by side (objects under lib/obj, git-ignored), then one link
        obj_dir = LIB.parent / "obj"
        obj_dir.mkdir(exist_ok=True)
        jobs = []
        for src in srcs:
            obj = obj_dir / (src.stem + ".o")
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
            if verbose:
                print(" ".join(cmd))
            jobs.append((subprocess.Popen(cmd), cmd, obj))
        objs = []
        for proc, cmd, obj in jobs:
            if proc.wait() != 0:
                raise subprocess.CalledProcessError(proc.returncode, cmd)
            objs.append(str(obj))
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *objs]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    drv = CSRC / "host_driver.cpp"
    if drv.exists() and (force or _stale(DRIVER, [drv, LIB])):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", str(DRIVER), str(drv), "-L" + str(LIB.parent), "-lhashjoin_b200",
               "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
