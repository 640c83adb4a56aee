"""Build libjpeg_gpu.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m imagecodecs_b200.build [--force]

Five kernel specialisations are compiled as separate translation units in parallel, then
linked with the host API into imagecodecs_b200/libjpeg_gpu.so (static cudart, no torch).
nvcc cross-compiles without a GPU, so this also runs in the CPU-only dev container.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libjpeg_gpu.so")

SPECS = [(0, 3), (0, 4), (1, 3), (1, 4), (2, 1)]   # (layout, channels): 444/420/gray
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# --fmad=false: the reference's float math must not be contracted
NVCC_FLAGS = ["-std=c++17", "-O3", "--fmad=false", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
              "-I" + CSRC, "-I" + os.path.join(ROOT, "include")] + ARCH


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "jpeg_gpu.h"), __file__]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_variant(out, defs, obj_dir):
    """Kernel-variant experiments: build a second library with extra -D flags (tools/ab_run.sh)."""
    global OBJ, LIB, NVCC_FLAGS
    saved = (OBJ, LIB, NVCC_FLAGS)
    OBJ, LIB, NVCC_FLAGS = obj_dir, out, NVCC_FLAGS + list(defs)
    try:
        return build(force=True, example=False)
    finally:
        OBJ, LIB, NVCC_FLAGS = saved


def build(force=False, verbose=False, example=True):
    deps = _deps()
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for layout, nc in SPECS:
        obj = os.path.join(OBJ, "kernel_%d_%d.o" % (layout, nc))
        jobs.append((obj, [_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v", "-DJG_LAYOUT=%d" % layout, "-DJG_NC=%d" % nc,
                                                   "-c", os.path.join(CSRC, "jpeg_kernel_inst.cu"), "-o", obj]))
    jobs.append((os.path.join(OBJ, "jpeg_stuff.o"), [_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", os.path.join(CSRC, "jpeg_stuff.cu"),
                                                                 "-o", os.path.join(OBJ, "jpeg_stuff.o")]))
    jobs.append((os.path.join(OBJ, "jpeg_entropy.o"), [_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", os.path.join(CSRC, "jpeg_entropy.cu"),
                                                                   "-o", os.path.join(OBJ, "jpeg_entropy.o")]))
    for src in ("jpeg_gpu_api.cpp", "jpeg_host.cpp", "codecs_jpeg.cpp", "jpeg_decode_host.cpp", "jpeg_decode_api.cpp"):
        if not os.path.exists(os.path.join(CSRC, src)):
            continue
        obj = os.path.join(OBJ, src.replace(".cpp", ".o"))
        jobs.append((obj, [_nvcc()] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]))
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        futs = {ex.submit(_run, cmd): obj for obj, cmd in jobs if force or _stale(obj, deps)}
        for f in cf.as_completed(futs):
            logs.append((futs[f], f.result()))
    # the reference's float math must never be contracted: no fused multiply-add of any width may
    # appear in the encode kernels (ptxas fuses packed mul+add despite .rn -- see jpeg_device.h)
    cuobjdump = os.path.join(os.path.dirname(_nvcc()), "cuobjdump")
    for obj in futs.values():
        if os.path.basename(obj).startswith("kernel_"):
            sass = _run([cuobjdump, "-sass", obj])
            fused = [l.strip() for l in sass.split("\n") if "FFMA" in l]
            if fused:
                raise RuntimeError("%s contains fused multiply-adds (output would differ from jpeg_enc.h):\n%s" % (obj, "\n".join(fused[:5])))
    _run([_nvcc(), "-shared", "-cudart", "static"] + ARCH + ["-o", LIB] + [obj for obj, _ in jobs])
    with open(os.path.join(OBJ, "ptxas.log"), "w") as fh:
        for obj, log in sorted(logs):
            fh.write("== %s\n%s\n" % (os.path.basename(obj), log))
    if not example:
        return LIB
    # the C++ example that drives the host layer the way the reference's tests.cpp does
    exe = os.path.join(HERE, "write_jpg_like_reference")
    _run(["g++", "-std=c++17", "-O2", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
          os.path.join(ROOT, "tests", "cpp", "write_jpg_like_reference.cpp"), "-o", exe,
          "-L" + HERE, "-ljpeg_gpu", "-Wl,-rpath,$ORIGIN"])
    if verbose:
        for obj, log in sorted(logs):
            print("==", os.path.basename(obj)); print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
