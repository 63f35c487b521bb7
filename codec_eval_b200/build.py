"""Builds libce_gpu.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libce_gpu.so")
SOURCES = ["ce_api.cu", "k_color.cu", "k_ssim2.cu", "k_dssim.cu", "k_butteraugli.cu", "k_jpeg.cu", "k_icc.cu"]
NVCC = os.environ.get("CE_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # explicit IEEE op sequences; fused multiply-adds are written as __fmaf_rn
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-shared", "-cudart", "static",
]


def built_hash() -> str:
    """The source hash stamped into the existing library's ce_version() ('' if there is none / it cannot be read)."""
    if not os.path.exists(OUT):
        return ""
    try:   # a child process: a stale library must not stay mapped in this one
        out = subprocess.run([sys.executable, "-c",
                              "import ctypes,sys; L=ctypes.CDLL(sys.argv[1]); L.ce_version.restype=ctypes.c_char_p; "
                              "print(L.ce_version().decode())", OUT], capture_output=True, text=True, timeout=120)
        v = out.stdout.strip()
        return v.rsplit("src:", 1)[-1] if "src:" in v else ""
    except Exception:
        return ""


def needs_build() -> bool:
    """True unless the library on disk was built from exactly the sources in the tree (content hash, not mtime)."""
    from . import _lib

    return built_hash() != _lib.source_hash()


def build(force: bool = False, verbose: bool = False, out: str = None, extra: list = None) -> str:
    """out / extra: an experiment build (other -D flags) written elsewhere; load it with CE_LIB_PATH=<out>."""
    if out is None and not force and not needs_build():
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError(f"{NVCC} not found: libce_gpu.so cannot be (re)built here and codec_eval_b200 has no CPU fallback")
    from . import _lib

    stamp = ["-DCE_SOURCE_HASH=\"" + _lib.source_hash() + ("+" + ",".join(extra) if extra else "") + "\""] + list(extra or [])
    tag = ("_" + "".join(ch if ch.isalnum() else "_" for ch in ",".join(extra))) if extra else ""
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", tag + ".o"))
        cmd = [NVCC] + [f for f in FLAGS if f not in ("-shared",)] + stamp + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {s}:\n{log}\n")
        elif verbose or log.strip():
            sys.stderr.write(log)
    if failed:
        raise RuntimeError("nvcc build failed")
    dst = out or OUT
    tmp = dst + f".{os.getpid()}.tmp"
    subprocess.check_call([NVCC] + FLAGS + objs + ["-o", tmp])
    os.replace(tmp, dst)
    return dst


def ensure_built(verbose: bool = False) -> str:
    """Build unless the library on disk already carries the hash of the sources in the tree.  Safe to call from several
    ranks at once (one builds under a file lock, the others wait and find it fresh)."""
    import fcntl

    if not needs_build():
        return OUT
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if needs_build():
                sys.stderr.write("codec_eval_b200: libce_gpu.so does not match the sources in the tree -- rebuilding (nvcc, sm_100a)\n")
                build(force=True, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    _extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out, extra=_extra))
