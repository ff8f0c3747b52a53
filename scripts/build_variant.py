"""Small experiment library: liblibint_b200<SUFFIX>.so holding only a few class kernels, compiled with extra
nvcc flags (kernel knobs are -D macros), linked against the default build's other objects.

    python scripts/build_variant.py _x1 2222,2122 -DLB200_SOME_KNOB=1

Use with LB200_LIB_SUFFIX=_x1 python scripts/prof_class.py 2 2 2 2 ...  (store-mode launches of those classes
only: the reduced dispatch table has no other class, so nothing that needs a Fock builder / Schwarz set-up)."""
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "libint_b200", "csrc")
LIBDIR = os.path.join(ROOT, "libint_b200", "_lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "--expt-relaxed-constexpr", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler",
         "-fvisibility=default", "-Xcompiler", "-fno-gnu-unique",   # per-library launcher statics: several variants share a process
         "-Wno-deprecated-gpu-targets", "-gencode", "arch=compute_100a,code=sm_100a"]


def main():
    suffix, classes = sys.argv[1], sys.argv[2].split(",")
    extra = sys.argv[3:]
    work = "/tmp/lb200_variant" + suffix
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(os.path.join(work, "gen"))
    with open(os.path.join(work, "gen", "dispatch_table.inc"), "w") as f:
        for c in classes:
            f.write("LB200_CLASS(%s, %s, %s, %s)\n" % tuple(c))
    with open(os.path.join(work, "inst.cu"), "w") as f:
        f.write('#include "launch.cuh"\nnamespace lb200 {\n')
        for c in classes:
            f.write("template cudaError_t launch_class_any<%s, %s, %s, %s>(const EriParams&, const RowInfo*, int, int, "
                    "cudaStream_t);\n" % tuple(c))
        f.write("}\n")
    shutil.copy(os.path.join(CSRC, "dispatch.cu"), os.path.join(work, "dispatch.cu"))
    objs = []
    for src in ("inst.cu", "dispatch.cu"):
        o = os.path.join(work, src + ".o")
        cmd = [NVCC] + FLAGS + extra + ["-I", work, "-I", CSRC, "-Xptxas", "-v", "-c", os.path.join(work, src), "-o", o]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.exit(r.stdout)
        if src == "inst.cu":
            lines = r.stdout.splitlines()
            for i, ln in enumerate(lines):
                if "Compiling entry function" in ln and ("Lb0ELb0E" in ln or "Li0EE" in ln):
                    name = ln.split("'")[1]
                    print(name[:70], "|", lines[i + 2].strip(), "|", lines[i + 3].strip() if i + 3 < len(lines) else "")
        objs.append(o)
    other = [o for o in glob.glob(os.path.join(LIBDIR, "obj", "*.o"))
             if not os.path.basename(o).startswith(("eri_inst_", "dispatch."))]
    lib = os.path.join(LIBDIR, "liblibint_b200%s.so" % suffix)
    r = subprocess.run([NVCC, "-shared", "-o", lib] + objs + other +
                       ["-gencode", "arch=compute_100a,code=sm_100a", "-Wno-deprecated-gpu-targets", "-ldl"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.exit(r.stdout)
    print(lib, "%.1f MB" % (os.path.getsize(lib) / 1e6))


if __name__ == "__main__":
    main()
