"""In-tree build of libwbem.so (sm_100a only) with nvcc.  `python -m wavebem_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
LIB = os.path.join(OUT, "libwbem.so")
SOURCES = ["api.cu", "assemble.cu", "operator.cu", "gmres.cu", "precond.cu", "spai.cu", "constraints.cu", "postproc.cu", "plan.cpp",
           "quadrature.cpp", "comm.cpp", "domain.cpp", "group.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    headers = [os.path.join(CSRC, "internal.h"), os.path.join(CSRC, "q1map.cuh"),
               os.path.join(HERE, "..", "include", "wbem.h")]
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + ARCH + FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(run, jobs):
            if (verbose or ptxas_info) and out:
                print(out)
    if force or jobs or _stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-ccbin", HOST_CXX, "-o", LIB] + objs + ["-ldl"])
    return LIB


def build_variant(tag: str, defines: list[str], sources=("assemble.cu",)) -> str:
    """Development helper: lib/libwbem_<tag>.so with `sources` recompiled with extra -D flags (kernel
    tuning experiments; the tile macros live in assemble.cu only) and the other objects of the main
    build; select it with the environment variable WBEM_LIB."""
    build()
    vdir = os.path.join(OUT, "variant_" + tag)
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for s in SOURCES:
        if s in sources:
            obj = os.path.join(vdir, os.path.splitext(s)[0] + ".o")
            subprocess.check_call([NVCC] + ARCH + FLAGS + defines + ["-c", os.path.join(CSRC, s), "-o", obj])
        else:
            obj = os.path.join(OUT, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
    lib = os.path.join(OUT, f"libwbem_{tag}.so")
    subprocess.check_call([NVCC] + ARCH + ["-shared", "-ccbin", HOST_CXX, "-o", lib] + objs + ["-ldl"])
    return lib


def build_probes() -> str:
    """lib/libwbem_probes.so: the FP64 issue-path micro-benchmarks (csrc/probes.cu) -- development only,
    never loaded by the product."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libwbem_probes.so")
    src = os.path.join(CSRC, "probes.cu")
    if _stale(lib, [src]):
        subprocess.check_call([NVCC] + ARCH + FLAGS + ["-shared", src, "-o", lib])
    return lib


def build_cpp_test() -> str:
    """Compile tests/cpp/host_mirror_test.cc (the C++ host mirror of BEMProblem<3>) against libwbem.so."""
    build()
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "cpp", "host_mirror_test.cc")
    exe = os.path.join(OUT, "host_mirror_test")
    hdr = os.path.join(CSRC, "bem_problem_b200.h")
    if _stale(exe, [src, hdr, LIB]):
        subprocess.check_call([HOST_CXX, "-O2", "-std=c++17", src, "-o", exe, "-L", OUT, "-lwbem",
                               "-Wl,-rpath,$ORIGIN"])
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
    if "--probes" in sys.argv:
        print(build_probes())
