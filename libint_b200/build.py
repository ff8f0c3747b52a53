"""Builds libint_b200/_lib/liblibint_b200.so in-tree with nvcc for sm_100a.

  python -m libint_b200.build [-j N] [--force]

The shared library is plain CUDA runtime + C ABI (include/libint_b200.h); it does not link
against torch.  Object files are cached under libint_b200/_lib/obj and rebuilt when their
sources (or any header) are newer.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment knobs: LB200_LIB_SUFFIX builds a variant library next to the default one,
# LB200_EXTRA_FLAGS appends nvcc flags (e.g. -DLB200_RR_MINB=3)
SUFFIX = os.environ.get("LB200_LIB_SUFFIX", "")
OUT = os.path.join(HERE, "_lib")
OBJ = os.path.join(OUT, "obj" + SUFFIX)
LIB = os.path.join(OUT, "liblibint_b200%s.so" % SUFFIX)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "--expt-relaxed-constexpr", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default", "-Wno-deprecated-gpu-targets"] + ARCH + \
    os.environ.get("LB200_EXTRA_FLAGS", "").split()
BOYS = os.path.join(HERE, "data", "boys_cheb7_m24.bin")


def _newer(src, dst, deps):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in [src] + deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(jobs=None, force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    gen = os.path.join(CSRC, "gen")
    if not os.path.isdir(gen) or not os.listdir(gen):
        _run([sys.executable, os.path.join(HERE, "tools", "gen_instances.py")])
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(gen, "*.inc")) + [os.path.join(HERE, "..", "include", "libint_b200.h")]
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(gen, "*.cu")))
    cc = sorted(glob.glob(os.path.join(CSRC, "*.cc")))
    jobs_list = []
    objs = []
    for s in cu + cc:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer(s, o, headers):
            jobs_list.append([NVCC] + NVCC_FLAGS + ["-I", CSRC, "-c", s, "-o", o])
    o = os.path.join(OBJ, "boys_table.o")
    objs.append(o)
    sfile = os.path.join(CSRC, "boys_table.S")
    if force or _newer(sfile, o, [BOYS]):
        jobs_list.append(["gcc", "-c", sfile, '-DBOYS_TABLE_PATH="%s"' % BOYS, "-o", o])
    jobs = jobs or min(8, os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for out in ex.map(_run, jobs_list):
            if verbose and out.strip():
                print(out)
    if jobs_list or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-Wno-deprecated-gpu-targets", "-ldl"])
    build_iface(force=bool(jobs_list) or force)
    build_driver(force=bool(jobs_list) or force)
    return LIB


IFACE_LIB = os.path.join(OUT, "liblibint_b200_iface%s.so" % SUFFIX)


def build_iface(force=False):
    """liblibint_b200_iface.so: the reference's Libint_t / libint2_build_* C boundary
    (csrc/iface/libint2_b200_iface.cc, plain host C++) on top of the CUDA library."""
    src = os.path.join(CSRC, "iface", "libint2_b200_iface.cc")
    inc = os.path.join(HERE, "..", "include")
    deps = glob.glob(os.path.join(inc, "libint2", "util", "generated", "*.h")) + \
        [os.path.join(inc, "libint_b200.h")]
    if force or _newer(src, IFACE_LIB, deps):
        _run([os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-fPIC", "-shared", "-I", inc, src, "-o", IFACE_LIB,
              "-L", OUT, "-l:" + os.path.basename(LIB), "-Wl,-rpath,$ORIGIN", "-lpthread",
              # link libstdc++ dynamically (this image's g++ wrapper otherwise finds only the static
              # archive, and a second copy inside the .so clashes with the host process's)
              "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-Wl,--exclude-libs,ALL"])
    return IFACE_LIB


DRIVER = os.path.join(OUT, "hartree-fock-b200%s" % SUFFIX)


def build_driver(force=False):
    """hartree-fock-b200: the reference's direct-SCF test driver as a C++ host program on the C ABI
    (csrc/tools/hartree_fock_b200.cc above include/libint_b200.hpp; plain g++, no CUDA headers)."""
    src = os.path.join(CSRC, "tools", "hartree_fock_b200.cc")
    inc = os.path.join(HERE, "..", "include")
    deps = [os.path.join(inc, "libint_b200.h"), os.path.join(inc, "libint_b200.hpp"), os.path.join(inc, "libint_b200_basis.hpp")]
    if force or _newer(src, DRIVER, deps):
        _run([os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-I", inc,
              '-DLIBINT_B200_DATADIR="%s"' % os.path.join(HERE, "data"), src, "-o", DRIVER, "-L", OUT,
              "-l:" + os.path.basename(LIB), "-Wl,-rpath,$ORIGIN"])
    return DRIVER


if __name__ == "__main__":
    j = None
    if "-j" in sys.argv:
        j = int(sys.argv[sys.argv.index("-j") + 1])
    print(build(jobs=j, force="--force" in sys.argv, verbose=True))
