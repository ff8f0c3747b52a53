#!/usr/bin/env python3
"""oracle/make_shims.py -- TEST INFRASTRUCTURE (oracle), not product code.

Prepares oracle/_ref/include/ so that the reference's own headers under
/root/reference/include compile unmodified where they lie:

  * libint2/config.h   <- /root/reference/include/libint2/config.h.in with the
                          `#undef X` lines the configure script would fill
                          replaced by the fixed oracle configuration below
  * libint2/basis.h    <- basis.h.in verbatim (it has no @...@ substitutions
                          that matter; the data path comes from LIBINT_DATA_PATH)
  * libint2/boost/     <- external/boost.tar.gz (Boost.Preprocessor only)

Nothing is written outside oracle/_ref/ (git-ignored build output).
"""
import os, re, sys, tarfile

REF = os.environ.get("LIBINT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "include", "libint2")

CONFIG = {
    "LIBINT_HOST_ARCH": '"x86_64"', "LIBINT_TARGET_ARCH": '"x86_64"',
    "LIBINT_VERSION": '"2.9.0"', "LIBINT_MAJOR_VERSION": "2",
    "LIBINT_MINOR_VERSION": "9", "LIBINT_MICRO_VERSION": "0",
    "LIBINT_API_PREFIX": '""', "LIBINT_HARD_MAX_AM": "10",
    "LIBINT_MAX_AM": "6", "LIBINT_OPT_AM": "3",
    "INCLUDE_ONEBODY": "0", "INCLUDE_ERI": "0", "INCLUDE_ERI3": "0", "INCLUDE_ERI2": "0",
    "MULTIPOLE_MAX_ORDER": "4", "ERI_MAX_AM": "6", "ERI3_MAX_AM": "6", "ERI2_MAX_AM": "6",
    "LIBINT_ENABLE_GENERIC_CODE": "1", "LIBINT_VECTOR_LENGTH": "1",
    "LIBINT_CGSHELL_ORDERING": "1", "LIBINT_SHGSHELL_ORDERING": "1", "LIBINT_SHELL_SET": "1",
    "LIBINT_CONTRACTED_INTS": "1", "LIBINT_SINGLE_EVALTYPE": "1", "LIBINT_ERI_STRATEGY": "1",
    "LIBINT_HAS_CXX11": "1", "LIBINT_HAS_EIGEN": "1",
    "LIBINT_CGSHELL_ORDERING_STANDARD": "1", "LIBINT_CGSHELL_ORDERING_INTV3": "2",
    "LIBINT_CGSHELL_ORDERING_GAMESS": "3", "LIBINT_CGSHELL_ORDERING_ORCA": "4",
    "LIBINT_CGSHELL_ORDERING_BAGEL": "5",
    "LIBINT_SHGSHELL_ORDERING_STANDARD": "1", "LIBINT_SHGSHELL_ORDERING_GAUSSIAN": "2",
    "LIBINT_SHELL_SET_STANDARD": "1", "LIBINT_SHELL_SET_ORCA": "2",
    "HAVE_STDINT_H": "1", "HAVE_POSIX_MEMALIGN": "1",
}


def main():
    os.makedirs(OUT, exist_ok=True)
    src = open(os.path.join(REF, "include/libint2/config.h.in")).read()

    def sub(m):
        name = m.group(1)
        return f"#define {name} {CONFIG[name]}" if name in CONFIG else f"/* #undef {name} */"

    cfg = re.sub(r"^#undef\s+(\w+)\s*$", sub, src, flags=re.M)
    open(os.path.join(OUT, "config.h"), "w").write(cfg)
    open(os.path.join(OUT, "basis.h"), "w").write(
        open(os.path.join(REF, "include/libint2/basis.h.in")).read())
    if not os.path.isdir(os.path.join(OUT, "boost")):
        with tarfile.open(os.path.join(REF, "external/boost.tar.gz")) as t:
            t.extractall(OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
