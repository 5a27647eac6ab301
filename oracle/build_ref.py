#!/usr/bin/env python
"""Compile the reference's OWN operator sources, where they lie under /root/reference,
into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  No reference source is copied
into this repository: g++/nvcc read the files in place and only the resulting .so files
land in oracle/_ref/.  We do not run the reference's build system (its "build system" is
a JIT `torch.utils.cpp_extension.load` at import time, cogdl/operators/spmm.py:11-40);
the commands below are our own short recipe with the same flags that JIT would use:

  CPU  : g++ -fPIC -std=c++17 -fopenmp            (NO -O flag: "as shipped")   -> _ref/asis/
         g++ -fPIC -std=c++17 -fopenmp -O3         ("fair" build)               -> _ref/o3/
  CUDA : nvcc -gencode arch=compute_100a,code=sm_100a -O3 (kernels unchanged)   -> _ref/cuda/

Modules (pybind names are hard-coded in the reference sources):
  spmm_cpu   cogdl/operators/spmm/spmm_cpu.cpp                      (the CPU oracle)
  sampler    cogdl/operators/sample/sample.cpp                      (coo2csr_cpu_index)
  spmm, sddmm, mhspmm, mhsddmm, mhtranspose, edge_softmax, scatter_max   (reference CUDA kernels:
             a second, GPU-side oracle, only executed by `-m gpu` tests on the B200 box)
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COGDL_REFERENCE", "/root/reference")
OPS = os.path.join(REF, "cogdl", "operators")
OUT = os.path.join(HERE, "_ref")

CPU_MODULES = {
    "spmm_cpu": ["spmm/spmm_cpu.cpp"],
    "sampler": ["sample/sample.cpp"],
}
CUDA_MODULES = {
    "spmm": (["spmm/spmm.cpp", "spmm/spmm_kernel.cu"], ["-lcusparse"]),
    "sddmm": (["spmm/sddmm.cpp", "spmm/sddmm_kernel.cu"], []),
    "mhspmm": (["spmm/multiheadSpmm.cpp", "spmm/multiheadSpmm.cu"], []),
    "mhsddmm": (["spmm/multiheadSddmm.cpp", "spmm/multiheadSddmm.cu"], []),
    "mhtranspose": (["spmm/mhTranspose.cpp", "spmm/mhTranspose.cu"], ["-lcusparse"]),
    "edge_softmax": (["edge_softmax/edge_softmax.cc", "edge_softmax/edge_softmax.cu"], []),
    "scatter_max": (["scatter_max/scatter_max.cc", "scatter_max/scatter_max.cu"], []),
}


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce

    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    defs = ["-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1"]
    link = [f"-L{libdir}", f"-Wl,-rpath,{libdir}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    return inc, defs, link, libdir


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-2000:] + r.stderr[-4000:] + "\n")
        raise RuntimeError("reference build failed: " + cmd[-1])


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cpu(name, variant, opt):
    inc, defs, link, _ = _torch_flags()
    srcs = [os.path.join(OPS, s) for s in CPU_MODULES[name]]
    outdir = os.path.join(OUT, variant)
    os.makedirs(outdir, exist_ok=True)
    target = os.path.join(outdir, name + ".so")
    if not _stale(target, srcs):
        return target
    cmd = ["g++", "-shared", "-fPIC", "-std=c++17", "-fopenmp", f"-DTORCH_EXTENSION_NAME={name}"]
    cmd += opt + defs + inc + srcs + ["-o", target] + link
    _run(cmd)
    return target


def build_cuda(name):
    inc, defs, link, libdir = _torch_flags()
    srcs, extra = CUDA_MODULES[name]
    srcs = [os.path.join(OPS, s) for s in srcs]
    outdir = os.path.join(OUT, "cuda")
    os.makedirs(outdir, exist_ok=True)
    target = os.path.join(outdir, name + ".so")
    if not _stale(target, srcs):
        return target
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [
        os.path.join(cuda_home, "bin", "nvcc"), "-shared", "-std=c++17", "-O3",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC", f"-DTORCH_EXTENSION_NAME={name}",
        "-I" + os.path.join(OPS, "spmm"),
    ]
    cmd += defs + inc + srcs + ["-o", target]
    cmd += [f"-L{libdir}", f"-Xlinker=-rpath,{libdir}", "-lc10", "-ltorch_cpu", "-ltorch",
            "-ltorch_python", "-lc10_cuda", "-ltorch_cuda"] + extra
    _run(cmd)
    return target


def main(cuda=True):
    if not os.path.isdir(OPS):
        print(f"[build_ref] {OPS} not present: keeping whatever is prebuilt in {OUT}")
        return
    jobs = []
    with ThreadPoolExecutor(max_workers=8) as ex:
        for name in CPU_MODULES:
            jobs.append(ex.submit(build_cpu, name, "asis", []))
            jobs.append(ex.submit(build_cpu, name, "o3", ["-O3"]))
        if cuda:
            for name in CUDA_MODULES:
                jobs.append(ex.submit(build_cuda, name))
        for j in jobs:
            print("[build_ref]", j.result())


if __name__ == "__main__":
    main(cuda="--no-cuda" not in sys.argv)
