"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

  librmd_b200.so   CUDA kernels + C ABI (include/rmd_b200.h), sm_100a only
  librmd_synth.so  host-only synthetic G-buffer generator (workload definition)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GCC = "/usr/bin/gcc"
CUDA_SOURCES = ["box_filter.cu", "weighted_filter.cu", "svgf_temporal.cu", "svgf_variance.cu", "svgf_atrous.cu", "svgf_ctx.cu", "p2p.cu"]
# svgf_atrous_tile.cu is compiled once per kernel variant (same list as RMD_ATROUS_VARIANTS in csrc/svgf.cuh)
ATROUS_VARIANTS = [0, 1, 3, 6, 7, 8, 11, 12, 15, 16]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
] + os.environ.get("RMD_EXTRA_NVCC", "").split()


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=False):
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")]
    headers.append(os.path.join(HERE, "..", "include", "rmd_b200.h"))
    headers.append(os.path.join(HERE, "..", "include", "rmd_b200_debug.h"))
    lib = os.path.join(HERE, "librmd_b200.so")
    jobs = []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or not _newer(o, [s] + headers):
            jobs.append((NVCC_FLAGS, s, o))
    tile_src = os.path.join(CSRC, "svgf_atrous_tile.cu")
    variant_objs = []
    for v in ATROUS_VARIANTS:
        o = os.path.join(objdir, f"svgf_atrous_tile_v{v}.o")
        variant_objs.append(o)
        if force or not _newer(o, [tile_src] + headers):
            jobs.append((NVCC_FLAGS + [f"-DRMD_VARIANT={v}"], tile_src, o))

    def compile_one(job):
        flags, s, o = job
        return _run([NVCC] + flags + ["-c", s, "-o", o], log=o + ".log")
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(compile_one, jobs))
    if verbose:
        for o in outs:
            print(o)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in CUDA_SOURCES] + variant_objs
    if force or jobs or not os.path.exists(lib):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    build_compat(force, bool(jobs))
    synth_src = os.path.join(HERE, "synth", "synth_scene.c")
    synth_lib = os.path.join(HERE, "librmd_synth.so")
    if force or not _newer(synth_lib, [synth_src]):
        _run([GCC, "-O3", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", synth_lib, synth_src, "-lm"])
    return lib, synth_lib


def build_compat(force=False, lib_rebuilt=False):
    """librmd_compat.a: the reference's caller-launched entry points (`filterKernelBaseline/Tiled<<<>>>`) as relocatable
    device code + the host classes of include/compat (Image, CudaGBuffer); and examples/build/compat_check, a program
    that uses them the way the reference's sources do."""
    comp = os.path.join(CSRC, "compat")
    inc = os.path.join(HERE, "..", "include")
    objdir = os.path.join(CSRC, "build")
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-rdc=true", "-Xcompiler", "-fPIC",
             "-I", os.path.join(inc, "compat"), "-I", inc]
    deps = [os.path.join(inc, "compat", h) for h in os.listdir(os.path.join(inc, "compat"))] + \
           [os.path.join(CSRC, "weighted.cuh"), os.path.join(inc, "rmd_b200.h")]
    objs, rebuilt = [], False
    for src in ("filter_compat.cu", "image.cpp", "gbuffer.cpp", "utils.cpp"):
        s = os.path.join(comp, src)
        o = os.path.join(objdir, "compat_" + src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + deps):
            _run([NVCC] + flags + ["-x", "cu", "-c", s, "-o", o], log=o + ".log")
            rebuilt = True
    lib = os.path.join(HERE, "librmd_compat.a")
    if rebuilt or not os.path.exists(lib):
        if os.path.exists(lib):
            os.remove(lib)
        _run(["/usr/bin/ar", "rcs", lib] + objs)
    exdir = os.path.join(HERE, "..", "examples", "build")
    os.makedirs(exdir, exist_ok=True)
    exe = os.path.join(exdir, "compat_check")
    src = os.path.join(HERE, "..", "examples", "compat_check.cu")
    if force or rebuilt or lib_rebuilt or not _newer(exe, [src, lib]):
        _run([NVCC] + flags + [src, lib, "-L", HERE, "-lrmd_b200", "-lz", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../raymarchdenoisercuda_b200",
                               "-o", exe])
    return lib, exe


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built")
