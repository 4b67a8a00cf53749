"""
Build recipe for the CPU oracle and (when the reference checkout is present) the reference
itself.  TEST INFRASTRUCTURE ONLY.

  oracle/libplatypus_oracle.so   our C restatement (oracle/platypus_oracle.c)
  oracle/_ref/libalign_ref.so    the UNMODIFIED reference kernel, compiled from where it lies:
                                 /root/reference/src/c/align.c (L1 ground truth)
  oracle/_ref/calign*.so         the reference's src/cython/calign.pyx compiled in a scratch
  oracle/_ref/calign_ref_wrap*.so   directory with three non-algorithmic accommodations
                                  + our forwarding wrapper (L2 ground truth)

Reference sources are never copied into the repo: scratch copies live in a temp dir that is
deleted afterwards, only binaries land in oracle/_ref/ (git-ignored, travels to the GPU box).
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLATYPUS_REFERENCE", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "libplatypus_oracle.so")
ALIGN_REF_SO = os.path.join(REF_OUT, "libalign_ref.so")
# the reference's own C flags, src/setup.py:31
REF_CFLAGS = ["-O2", "-funroll-loops", "-fPIC"]


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    src = [os.path.join(HERE, "platypus_oracle.c"), os.path.join(HERE, "platypus_oracle.h"),
           os.path.join(HERE, "..", "include", "platypus_b200.h")]
    if not force and _newer(ORACLE_SO, src):
        return ORACLE_SO
    _run(["gcc", "-O2", "-funroll-loops", "-fPIC", "-fopenmp", "-shared", "-std=c11", src[0], "-o", ORACLE_SO, "-lm"])
    return ORACLE_SO


def have_reference():
    return os.path.exists(os.path.join(REF, "src", "c", "align.c"))


def build_align_ref(force=False):
    """L1: gcc on the reference's own file, flags of src/setup.py:31."""
    if not have_reference():
        return ALIGN_REF_SO if os.path.exists(ALIGN_REF_SO) else None
    src = os.path.join(REF, "src", "c", "align.c")
    if not force and _newer(ALIGN_REF_SO, [src]):
        return ALIGN_REF_SO
    os.makedirs(REF_OUT, exist_ok=True)
    _run(["gcc"] + REF_CFLAGS + ["-shared", "-I" + os.path.join(REF, "src", "c"), src, "-o", ALIGN_REF_SO])
    return ALIGN_REF_SO


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def calign_ref_paths():
    return (os.path.join(REF_OUT, "calign" + _ext_suffix()),
            os.path.join(REF_OUT, "calign_ref_wrap" + _ext_suffix()))


def build_calign_ref(force=False):
    """L2: the reference's calign.pyx, cythonized in a scratch dir.

    Accommodations, none touching the algorithm:
      (i)   calign.pyx:26  `4**hash_nucs` -> `16384` (Cython 3 types ** as double)
      (ii)  calign.pyx:11  runtime `import htslibWrapper` commented out (no htslib here)
      (iii) htslibWrapper.pxd replaced by an excerpt of itself: the cAlignedRead struct
            (lines 187-201) and the flag DEFs / inline accessors (233-296)
    """
    mod, wrap = calign_ref_paths()
    if not have_reference():
        return (mod, wrap) if os.path.exists(mod) and os.path.exists(wrap) else None
    try:
        import Cython  # noqa: F401
    except ImportError:
        return None
    srcs = [os.path.join(REF, "src", "cython", "calign.pyx"), os.path.join(HERE, "calign_ref_wrap.pyx")]
    if not force and _newer(mod, srcs) and _newer(wrap, srcs):
        return mod, wrap
    os.makedirs(REF_OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="plb_calign_")
    try:
        cy = os.path.join(REF, "src", "cython")
        for f in ("calign.pyx", "calign.pxd", "cerrormodel.pxd"):
            shutil.copy(os.path.join(cy, f), tmp)
        shutil.copy(os.path.join(HERE, "calign_ref_wrap.pyx"), tmp)
        p = os.path.join(tmp, "calign.pyx")
        s = open(p).read()
        assert "cdef int hash_size = 4**hash_nucs" in s and "\nimport htslibWrapper\n" in s
        s = s.replace("cdef int hash_size = 4**hash_nucs", "cdef int hash_size = 16384")
        s = s.replace("\nimport htslibWrapper\n", "\n#import htslibWrapper\n")
        open(p, "w").write(s)
        lines = open(os.path.join(cy, "htslibWrapper.pxd")).read().split("\n")
        excerpt = lines[186:201] + [""] + lines[232:296]
        assert excerpt[0].startswith("ctypedef struct cAlignedRead")
        open(os.path.join(tmp, "htslibWrapper.pxd"), "w").write("\n".join(excerpt) + "\n")
        inc = sysconfig.get_paths()["include"]
        _run([sys.executable, "-m", "cython", "-2", "-I", tmp, "calign.pyx", "-o", "calign.c"], cwd=tmp)
        _run([sys.executable, "-m", "cython", "-3", "-I", tmp, "calign_ref_wrap.pyx", "-o", "calign_ref_wrap.c"], cwd=tmp)
        cflags = REF_CFLAGS + ["-shared", "-w", "-I" + tmp, "-I" + inc, "-I" + os.path.join(REF, "src", "c")]
        _run(["gcc"] + cflags + ["calign.c", os.path.join(REF, "src", "c", "align.c"), "-o", mod], cwd=tmp)
        _run(["gcc"] + cflags + ["calign_ref_wrap.c", "-o", wrap], cwd=tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return mod, wrap


def build_all(force=False, verbose=False):
    out = {"oracle": build_oracle(force), "align_ref": build_align_ref(force)}
    try:
        out["calign_ref"] = build_calign_ref(force)
    except Exception as e:  # the L2 reference build is a bonus; L1 + restatement still stand
        out["calign_ref"] = None
        out["calign_ref_error"] = str(e)
    if verbose:
        for k, v in out.items():
            print(k, "->", v)
    return out


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
