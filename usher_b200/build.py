"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo to the GPU box).

  usher_b200/libusher_b200.so   C-ABI + sm_100a kernels (nvcc -gencode arch=compute_100a,code=sm_100a)
  usher_b200/libub200_synth.so  synthetic MAT / sample generator (host only)
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
PROFILE = bool(os.environ.get("UB200_PROFILE"))   # developer build with in-kernel cycle counters
# developer A/B builds of the kernel: UB200_VARIANT="acc2,nofence" compiles with -DUB200_V_ACC2 -DUB200_V_NOFENCE into
# libusher_b200_v_acc2_nofence.so (scripts/variants.py times them side by side)
VARIANT = [v for v in os.environ.get("UB200_VARIANT", "").split(",") if v]
LIB = os.path.join(PKG, "libusher_b200_prof.so" if PROFILE else
                   ("libusher_b200_v_" + "_".join(VARIANT) + ".so" if VARIANT else "libusher_b200.so"))
SYNTH = os.path.join(PKG, "libub200_synth.so")
USHER = os.path.join(PKG, "usher")   # the drop-in CLI (host C++ over the C ABI)
# The image exports CXX=/opt/gcc/bin/g++, a wrapper that links libstdc++ statically; a second libstdc++ in a
# python process that already loaded the shared one crashes.  Pin the system compiler.
HOSTCXX = "/usr/bin/g++"


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs += [os.path.join(INC, f) for f in os.listdir(INC)]
    srcs = [os.path.join(CSRC, "api.cu"), os.path.join(CSRC, "derive.cpp")]
    if force or _stale(LIB, srcs + hdrs):
        cmd = [
            _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC", "-shared", "-I", INC, "-I", CSRC,
            "-Xptxas", "-v" if verbose else "-warn-spills", "-o", LIB,
        ] + (["-DUB200_PROFILE"] if PROFILE else []) + ["-DUB200_V_" + v.upper() for v in VARIANT] + srcs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed")
    ssrc = [os.path.join(CSRC, "synth.cpp")]
    if force or _stale(SYNTH, ssrc + hdrs):
        subprocess.check_call([HOSTCXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-I", INC, "-o", SYNTH] + ssrc)
    hdir = os.path.join(CSRC, "host")
    hsrc = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir)) if f.endswith(".cpp")]
    hhdr = [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".hpp")]
    if not PROFILE and not VARIANT and (force or _stale(USHER, hsrc + hhdr + hdrs + [LIB])):
        subprocess.check_call([HOSTCXX, "-std=c++17", "-O2", "-I", INC, "-I", hdir] + hsrc +
                              ["-o", USHER, "-L", PKG, "-lusher_b200", "-Wl,-rpath,$ORIGIN", "-lz"])
    return LIB, SYNTH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
