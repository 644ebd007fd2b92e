"""Builds libplb200.so in-tree with nvcc for sm_100a (no torch in the link)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libplb200.so")
SOURCES = ["api.cu", "photo.cu", "photo_min.cu", "ssim.cu", "smooth.cu", "edge.cu", "warp.cu", "cloud.cu", "velo.cu", "prep.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "-I", os.path.join(ROOT, "include")]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "plb200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu to an object (parallel) and link the shared library."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("PLB_NVCC_EXTRA", "").split()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        objs.append(obj)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                 "-Xcompiler", "-fPIC"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + out.stdout)
    return LIB


TORCH_LIB = os.path.join(HERE, "_plb200_torch.so")


def build_torch(force=False):
    """The thin torch C++ binding (csrc/torch_binding.cpp): plain g++ against torch's headers, linked to
    libplb200.so beside it ($ORIGIN).  Host code only - the kernels live in libplb200.so."""
    src = os.path.join(CSRC, "torch_binding.cpp")
    deps = [src, os.path.join(ROOT, "include", "plb200.h"), __file__]
    if not force and os.path.isfile(TORCH_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(TORCH_LIB) for d in deps):
        return TORCH_LIB
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    tlib = ce.library_paths()[0]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", TORCH_LIB,
           "-DTORCH_EXTENSION_NAME=_plb200_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-I", os.path.join(ROOT, "include"), "-I", sysconfig.get_paths()["include"], "-I", os.path.join(cuda_home, "include")]
    for inc in ce.include_paths():
        cmd += ["-isystem", inc]
    cmd += ["-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-L", HERE, "-l:libplb200.so", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib, "-Wno-attributes"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("torch binding failed to build:\n" + out.stdout[-4000:])
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--torch" in sys.argv:
        print(build_torch(force="--force" in sys.argv))
