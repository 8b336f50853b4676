"""Build lib2019global_b200.so (the C-ABI engine) in-tree for sm_100a.

    python -m 2019global_b200.build        (or __graft_entry__.build())

Per translation unit:
  scene.cpp, builtin_scenes.cpp   host C++, -ffp-contract=off (bit-exact geometry)
  ref_kernels.cu                  -fmad=false (bit-exact reference arithmetic)
  path_kernels.cu                 --use_fast_math (FP32 tracer: MUFU rcp/rsqrt/sin/cos)
  path.cu, engine.cu
All with -gencode arch=compute_100a,code=sm_100a -lineinfo. nvcc cross-compiles
without a GPU; the resulting .so travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib2019global_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
          "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
          # hardened libstdc++ on the host side (bounds-checked operator[] etc.): every GPU test then also checks
          # the scene / tree builders for out-of-range accesses; the cost is invisible next to the kernels
          "-D_GLIBCXX_ASSERTIONS"]
UNITS = [
    ("scene.cpp", []),
    ("builtin_scenes.cpp", []),
    ("fixed_shapes.cpp", []),
    ("ref_kernels.cu", ["-fmad=false", "-Xptxas", "-v"]),
    ("path_kernels.cu", ["--use_fast_math", "-Xptxas", "-v"]),
    ("frame_kernels.cu", []),
    ("tree_build.cu", []),
    ("ray_sort.cu", []),
    ("bvh_build.cu", []),
    ("path.cu", []),
    ("engine.cu", []),
]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "g19.h"))
    objs = []
    logs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if not force and not _stale(o, [s] + headers):
            continue
        cmd = [NVCC] + ARCH + COMMON + extra + ["-x", "cu", "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append((src, r.stderr))
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on " + src)
    if force or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB, logs


if __name__ == "__main__":
    lib, logs = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(lib)
