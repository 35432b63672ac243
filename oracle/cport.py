"""ctypes binding of oracle/qc_oracle.c (the C restatement: checker + timed CPU baseline).

TEST INFRASTRUCTURE.  ``load()`` builds ``oracle/_build/libqc_oracle.so`` on demand with the system
gcc (portable ``-march=x86-64-v3`` so a library built in the CPU container also runs on the GPU box);
``load(native=True)`` builds a ``-march=native`` copy on the machine it runs on -- that is the one
``bench.py`` times, so the CPU baseline gets the host's full vector ISA.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_GCC = "/usr/bin/gcc"  # the image's $CC (/opt/gcc/bin/gcc) has no libgomp.spec
_libs: dict[bool, ctypes.CDLL] = {}


def build(native: bool = False, force: bool = False) -> Path:
    _BUILD.mkdir(exist_ok=True)
    name = "libqc_oracle_native.so" if native else "libqc_oracle.so"
    out = _BUILD / name
    src = _HERE / "qc_oracle.c"
    if out.exists() and not force and out.stat().st_mtime >= src.stat().st_mtime:
        return out
    march = "native" if native else "x86-64-v3"
    tmp = out.with_suffix(f".{os.getpid()}.tmp")
    # -ffp-contract=off: BEAM never fuses a*b+c, so the f64 restatements must not either (the timed f32
    # encoder re-enables contraction with a function attribute)
    cmd = [_GCC, "-O3", f"-march={march}", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", str(tmp),
           str(src), "-lm"]
    subprocess.run(cmd, check=True, capture_output=True)
    os.replace(tmp, out)
    return out


def load(native: bool = False) -> ctypes.CDLL:
    if native in _libs:
        return _libs[native]
    lib = ctypes.CDLL(str(build(native)))
    dp = ctypes.POINTER(ctypes.c_double)
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    up = ctypes.POINTER(ctypes.c_uint16)
    lib.qo_cosine_rules.restype = ctypes.c_double
    lib.qo_cosine_rules.argtypes = [dp, dp, ctypes.c_int]
    lib.qo_cosine_aggregator.restype = ctypes.c_double
    lib.qo_cosine_aggregator.argtypes = [dp, dp, ctypes.c_int]
    lib.qo_similarity_batch.restype = None
    lib.qo_similarity_batch.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int,
                                        ctypes.c_int, ip, ip, dp]
    lib.qo_chunk_mean.restype = None
    lib.qo_chunk_mean.argtypes = [dp, ctypes.c_int, ctypes.c_int, dp]
    lib.qo_first_match.restype = ctypes.c_int
    lib.qo_first_match.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    lib.qo_encode_f64.restype = None
    lib.qo_encode_f64.argtypes = [up, up, ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, dp]
    lib.qo_encode_f32.restype = None
    lib.qo_encode_f32.argtypes = [up, up, ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, fp]
    lib.qo_max_threads.restype = ctypes.c_int
    lib.qo_max_threads.argtypes = []
    _libs[native] = lib
    return lib


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def cosine_rules(a, b, native=False) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape
    return load(native).qo_cosine_rules(_p(a, ctypes.c_double), _p(b, ctypes.c_double), a.size)


def cosine_aggregator(a, b, native=False) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape
    return load(native).qo_cosine_aggregator(_p(a, ctypes.c_double), _p(b, ctypes.c_double), a.size)


def similarity_batch(emb, thr, want_full=True, threads=1, want_cos=True, native=False):
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    b, n, d = emb.shape
    thr = np.ascontiguousarray(thr, dtype=np.float64)
    ref_ok = np.zeros(b, dtype=np.int32)
    medoid = np.zeros(b, dtype=np.int32)
    cos = np.zeros((b, n, n), dtype=np.float64) if (want_cos and want_full) else None
    load(native).qo_similarity_batch(_p(emb, ctypes.c_float), b, n, d, _p(thr, ctypes.c_double),
                                     1 if want_full else 0, threads, _p(ref_ok, ctypes.c_int32),
                                     _p(medoid, ctypes.c_int32),
                                     _p(cos, ctypes.c_double) if cos is not None else None)
    return ref_ok, medoid, cos


def chunk_mean(chunks) -> np.ndarray:
    chunks = np.ascontiguousarray(chunks, dtype=np.float64)
    c, d = chunks.shape
    out = np.zeros(d, dtype=np.float64)
    load().qo_chunk_mean(_p(chunks, ctypes.c_double), c, d, _p(out, ctypes.c_double))
    return out


def first_match(query, rows, thr) -> int:
    query = np.ascontiguousarray(query, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    m, d = rows.shape
    return load().qo_first_match(_p(query, ctypes.c_double), _p(rows, ctypes.c_double), m, d, float(thr))


def encode(table_bf16, w_bf16, token_ids, lens, threads=1, f32=False, native=False) -> np.ndarray:
    table = np.ascontiguousarray(table_bf16, dtype=np.uint16)
    w = np.ascontiguousarray(w_bf16, dtype=np.uint16)
    vocab, dim = table.shape
    token_ids = np.ascontiguousarray(token_ids, dtype=np.int32)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    r, t = token_ids.shape
    lib = load(native)
    if f32:
        out = np.zeros((r, dim), dtype=np.float32)
        lib.qo_encode_f32(_p(table, ctypes.c_uint16), _p(w, ctypes.c_uint16), vocab, dim,
                          _p(token_ids, ctypes.c_int32), _p(lens, ctypes.c_int32), r, t, threads,
                          _p(out, ctypes.c_float))
    else:
        out = np.zeros((r, dim), dtype=np.float64)
        lib.qo_encode_f64(_p(table, ctypes.c_uint16), _p(w, ctypes.c_uint16), vocab, dim,
                          _p(token_ids, ctypes.c_int32), _p(lens, ctypes.c_int32), r, t, threads,
                          _p(out, ctypes.c_double))
    return out


def max_threads(native=False) -> int:
    return load(native).qo_max_threads()
