"""Builds quoracle_b200/lib/libqconsensus.so (sm_100a only) with nvcc.  Cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "lib" / "libqconsensus.so"
SOURCES = ["qc_api.cu", "qc_encoder.cu", "qc_similarity.cu", "qc_similarity_tc.cu"]
HEADERS = ["qc_common.cuh", "ptx.cuh", "../../include/qconsensus.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-ccbin", "/usr/bin/g++",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """One nvcc -c per source (in parallel; objects cached under lib/obj by mtime), then one link."""
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    obj_dir = LIB.parent / "obj"
    obj_dir.mkdir(exist_ok=True)
    hdr_t = max(d.stat().st_mtime for d in [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)])
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs = []
    for src in SOURCES:
        obj = obj_dir / (src + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(hdr_t, (CSRC / src).stat().st_mtime):
            continue
        cmd = [_nvcc(), *compile_flags, "-c", "-o", str(obj), str(CSRC / src)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out)
            failed = True
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libqconsensus.so")
    tmp = LIB.with_suffix(f".{os.getpid()}.tmp.so")
    link = [_nvcc(), "-shared", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
            *[str(obj_dir / (s + ".o")) for s in SOURCES]]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libqconsensus.so")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
