"""ctypes binding of libqconsensus.so -- the only way this package reaches the GPU.

There is no CPU fallback: if the shared library is missing or no sm_100 device is usable the calls
raise.  The binding mirrors ``include/qconsensus.h`` one to one; array arguments are numpy arrays
(host entry points) or integer device addresses (``*_dev`` entry points, e.g. ``tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

import os

# QCONSENSUS_LIB: tuning hook -- load an alternative build of the same library (tools/ only)
_LIB_PATH = Path(os.environ.get("QCONSENSUS_LIB") or Path(__file__).resolve().parent / "lib" / "libqconsensus.so")
_lib: Optional[ctypes.CDLL] = None

QC_K_NAMES = ("encode", "finalize", "similarity", "first_match", "chunk_mean")
QC_MAX_POOL = 64
QC_MAX_TOKENS = 8192
QC_PAD_TOKEN = -1

# Every symbol include/qconsensus.h declares (checked by tests/test_abi_symbols.py).
ABI_SYMBOLS = (
    "qc_ctx_create", "qc_ctx_destroy", "qc_ctx_device", "qc_ctx_dim", "qc_ctx_synchronize",
    "qc_host_alloc", "qc_host_free",
    "qc_decide_batch", "qc_embed_ragged", "qc_decide_ragged", "qc_decide_ragged_submit", "qc_decide_batch_submit", "qc_decide_batch_wait", "qc_embed_batch", "qc_similarity_batch", "qc_first_match_batch", "qc_chunk_mean",
    "qc_decide_batch_dev", "qc_embed_batch_dev", "qc_similarity_batch_dev",
    "qc_strerror", "qc_status_atom", "qc_abi_version",
    "qc_check_status", "qc_last_hang",
    "qc_launch_count", "qc_profile_enable", "qc_profile_reset", "qc_kernel_stats",
    "qc_encoder_info", "qc_encoder_set_cta_group", "qc_encoder_set_variant", "qc_encoder_stall_profile",
    "qc_similarity_set_path", "qc_encoder_set_multicast", "qc_encoder_multicast",
    "qc_capture_begin", "qc_capture_end", "qc_graph_launch", "qc_graph_destroy",
    "qc_lesson_cache_create", "qc_lesson_cache_destroy", "qc_lesson_cache_count", "qc_lesson_cache_put",
    "qc_lesson_cache_gather", "qc_lesson_cache_match", "qc_encoder_token_block", "qc_encoder_hint_block_map", "qc_encoder_set_packing", "qc_debug_pack_plan",
    "qc_split_create", "qc_split_export", "qc_split_attach", "qc_decide_split_dev", "qc_split_gathered",
    "qc_split_destroy", "qc_tsplit_create", "qc_embed_tsplit_dev", "qc_decide_tsplit_dev",
)


def debug_pack_plan(lens: np.ndarray, t: int, token_block: int):
    """Host-only (no GPU): the packed-stream plan the host entry points would build -> dict of int32 arrays."""
    lib = load_library()
    lens = np.ascontiguousarray(lens, dtype=np.int32).reshape(-1)
    r = lens.shape[0]
    cap = 3 * (int(lens.sum()) // 8 + r + token_block) + r + 1
    out = np.zeros(cap, np.int32)
    nchunks, nruns = ctypes.c_int32(), ctypes.c_int32()
    _check(lib.qc_debug_pack_plan(_ptr(lens), r, int(t), int(token_block), _ptr(out), cap, ctypes.byref(nchunks),
                                  ctypes.byref(nruns)))
    n = nchunks.value
    return {"chunk_src": out[:n], "chunk_n": out[n:2 * n], "chunk_row": out[2 * n:3 * n],
            "row_start": out[3 * n:3 * n + r + 1], "nruns": nruns.value}


class QcError(RuntimeError):
    """A negative qc_status; ``atom`` is the Elixir error atom the status maps to."""

    def __init__(self, status: int, message: str, atom: str):
        super().__init__(f"libqconsensus: {message} (status {status}, :{atom})")
        self.status = status
        self.atom = atom


class _EncCfg(ctypes.Structure):
    _fields_ = [("vocab", ctypes.c_int32), ("dim", ctypes.c_int32),
                ("table_bf16", ctypes.c_void_p), ("w_bf16", ctypes.c_void_p)]


def load_library() -> ctypes.CDLL:
    """Load the CUDA extension; raise loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m quoracle_b200.build` "
            "(or __graft_entry__.build()).  quoracle_b200 has no CPU fallback.")
    lib = ctypes.CDLL(str(_LIB_PATH))
    vp, i32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.qc_ctx_create.restype = i32
    lib.qc_ctx_create.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(_EncCfg)]
    lib.qc_ctx_destroy.restype = None
    lib.qc_ctx_destroy.argtypes = [vp]
    lib.qc_ctx_device.restype = i32
    lib.qc_ctx_device.argtypes = [vp]
    lib.qc_ctx_dim.restype = i32
    lib.qc_ctx_dim.argtypes = [vp]
    lib.qc_ctx_synchronize.restype = i32
    lib.qc_ctx_synchronize.argtypes = [vp]
    lib.qc_host_alloc.restype = vp
    lib.qc_host_alloc.argtypes = [sz]
    lib.qc_host_free.restype = None
    lib.qc_host_free.argtypes = [vp]
    lib.qc_decide_batch.restype = i32
    lib.qc_decide_batch.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.qc_decide_batch_submit.restype = i32
    lib.qc_decide_batch_submit.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.qc_decide_batch_wait.restype = i32
    lib.qc_decide_batch_wait.argtypes = [vp, i32]
    lib.qc_embed_batch.restype = i32
    lib.qc_embed_batch.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.qc_similarity_batch.restype = i32
    lib.qc_similarity_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.qc_first_match_batch.restype = i32
    lib.qc_first_match_batch.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.qc_chunk_mean.restype = i32
    lib.qc_chunk_mean.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.qc_decide_batch_dev.restype = i32
    lib.qc_decide_batch_dev.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.qc_embed_batch_dev.restype = i32
    lib.qc_embed_batch_dev.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.qc_similarity_batch_dev.restype = i32
    lib.qc_similarity_batch_dev.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.qc_embed_ragged.restype = i32
    lib.qc_embed_ragged.argtypes = [vp, vp, vp, i32, vp]
    lib.qc_decide_ragged_submit.restype = i32
    lib.qc_decide_ragged_submit.argtypes = [vp, i32, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.qc_decide_ragged.restype = i32
    lib.qc_decide_ragged.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.qc_split_create.restype = i32
    lib.qc_split_create.argtypes = [vp, i32, i32, i32, i32, ctypes.POINTER(vp)]
    lib.qc_split_export.restype = i32
    lib.qc_split_export.argtypes = [vp, ctypes.c_char_p]
    lib.qc_split_attach.restype = i32
    lib.qc_split_attach.argtypes = [vp, i32, ctypes.c_char_p]
    lib.qc_decide_split_dev.restype = i32
    lib.qc_decide_split_dev.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.qc_split_gathered.restype = i32
    lib.qc_split_gathered.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i32)]
    lib.qc_tsplit_create.restype = i32
    lib.qc_tsplit_create.argtypes = [vp, i32, i32, i32, ctypes.POINTER(vp)]
    lib.qc_embed_tsplit_dev.restype = i32
    lib.qc_embed_tsplit_dev.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.qc_decide_tsplit_dev.restype = i32
    lib.qc_decide_tsplit_dev.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.qc_split_destroy.restype = i32
    lib.qc_split_destroy.argtypes = [vp]
    lib.qc_strerror.restype = ctypes.c_char_p
    lib.qc_strerror.argtypes = [i32]
    lib.qc_status_atom.restype = ctypes.c_char_p
    lib.qc_status_atom.argtypes = [i32]
    lib.qc_abi_version.restype = i32
    lib.qc_abi_version.argtypes = []
    lib.qc_check_status.restype = i32
    lib.qc_check_status.argtypes = [vp]
    lib.qc_last_hang.restype = i32
    lib.qc_last_hang.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    lib.qc_launch_count.restype = ctypes.c_uint64
    lib.qc_launch_count.argtypes = [vp]
    lib.qc_profile_enable.restype = i32
    lib.qc_profile_enable.argtypes = [vp, i32]
    lib.qc_profile_reset.restype = i32
    lib.qc_profile_reset.argtypes = [vp]
    lib.qc_kernel_stats.restype = i32
    lib.qc_kernel_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    lib.qc_encoder_info.restype = i32
    lib.qc_encoder_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                    ctypes.POINTER(ctypes.c_int32)]
    lib.qc_encoder_set_variant.restype = i32
    lib.qc_encoder_set_variant.argtypes = [vp, i32, i32, i32, i32]
    lib.qc_encoder_set_multicast.restype = i32
    lib.qc_encoder_set_multicast.argtypes = [vp, i32]
    lib.qc_encoder_multicast.restype = i32
    lib.qc_encoder_multicast.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    lib.qc_capture_begin.restype = i32
    lib.qc_capture_begin.argtypes = [vp, vp]
    lib.qc_capture_end.restype = i32
    lib.qc_capture_end.argtypes = [vp, vp, ctypes.POINTER(vp)]
    lib.qc_graph_launch.restype = i32
    lib.qc_graph_launch.argtypes = [vp, vp]
    lib.qc_graph_destroy.restype = None
    lib.qc_graph_destroy.argtypes = [vp]
    lib.qc_lesson_cache_create.restype = i32
    lib.qc_lesson_cache_create.argtypes = [vp, i32, i32, ctypes.POINTER(vp)]
    lib.qc_lesson_cache_destroy.restype = None
    lib.qc_lesson_cache_destroy.argtypes = [vp]
    lib.qc_lesson_cache_count.restype = i32
    lib.qc_lesson_cache_count.argtypes = [vp]
    lib.qc_lesson_cache_put.restype = i32
    lib.qc_lesson_cache_put.argtypes = [vp, i32, vp, i32]
    lib.qc_lesson_cache_gather.restype = i32
    lib.qc_lesson_cache_gather.argtypes = [vp, vp, i32]
    lib.qc_lesson_cache_match.restype = i32
    lib.qc_lesson_cache_match.argtypes = [vp, vp, ctypes.c_float, ctypes.POINTER(ctypes.c_int32), vp]
    lib.qc_debug_pack_plan.restype = i32
    lib.qc_debug_pack_plan.argtypes = [vp, i32, i32, i32, vp, ctypes.c_int64, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.qc_encoder_set_packing.restype = i32
    lib.qc_encoder_set_packing.argtypes = [vp, i32]
    lib.qc_encoder_token_block.restype = i32
    lib.qc_encoder_token_block.argtypes = [vp, i32, ctypes.POINTER(i32)]
    lib.qc_encoder_hint_block_map.restype = i32
    lib.qc_encoder_hint_block_map.argtypes = [vp, vp, i32]
    lib.qc_similarity_set_path.restype = i32
    lib.qc_similarity_set_path.argtypes = [vp, i32]
    lib.qc_encoder_stall_profile.restype = i32
    lib.qc_encoder_stall_profile.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    lib.qc_encoder_set_cta_group.restype = i32
    lib.qc_encoder_set_cta_group.argtypes = [vp, i32]
    _lib = lib
    return lib


def _check(status: int) -> None:
    if status != 0:
        lib = load_library()
        raise QcError(status, lib.qc_strerror(status).decode(), lib.qc_status_atom(status).decode())


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _dev(addr) -> ctypes.c_void_p:
    return ctypes.c_void_p(int(addr) if addr else 0)


class PinnedArray:
    """numpy view over qc_host_alloc memory (page-locked, so the library copies from it directly)."""

    def __init__(self, shape, dtype):
        lib = load_library()
        self._dtype = np.dtype(dtype)
        self._nbytes = int(np.prod(shape)) * self._dtype.itemsize
        self._addr = lib.qc_host_alloc(max(self._nbytes, 1))
        if not self._addr:
            raise QcError(-5, "pinned allocation failed", "service_unavailable")
        buf = (ctypes.c_char * max(self._nbytes, 1)).from_address(self._addr)
        self.array = np.frombuffer(buf, dtype=self._dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._addr:
            self.array = None
            load_library().qc_host_free(ctypes.c_void_p(self._addr))
            self._addr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class DecideResult:
    ref_ok: np.ndarray            # [B] int32, reference rule verdict (value_0 wins when 1)
    medoid: np.ndarray            # [B] int32, north-star vote
    cos: Optional[np.ndarray]     # [B, N, N] float32
    emb: Optional[np.ndarray]     # [B, N, D] float32


class Context:
    """One libqconsensus context == one GPU (one process per GPU)."""

    def __init__(self, device: int = 0, table_bf16: Optional[np.ndarray] = None,
                 w_bf16: Optional[np.ndarray] = None):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        cfg = None
        if (table_bf16 is None) != (w_bf16 is None):
            raise ValueError("table_bf16 and w_bf16 must be given together")
        if table_bf16 is not None:
            table = np.ascontiguousarray(table_bf16, dtype=np.uint16)
            w = np.ascontiguousarray(w_bf16, dtype=np.uint16)
            if table.ndim != 2 or w.shape != (table.shape[1], table.shape[1]):
                raise ValueError("table must be [vocab, dim] and w [dim, dim] (bf16 bit patterns)")
            cfg = _EncCfg(table.shape[0], table.shape[1], table.ctypes.data, w.ctypes.data)
            self._keep = (table, w)
        _check(self._lib.qc_ctx_create(ctypes.byref(self._h), int(device),
                                       ctypes.byref(cfg) if cfg is not None else None))
        self.device = int(device)
        self.dim = int(self._lib.qc_ctx_dim(self._h))
        self._keep = None  # weights were copied to the device

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            self._lib.qc_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self) -> None:
        _check(self._lib.qc_ctx_synchronize(self._h))

    # -- host entry points ---------------------------------------------------------------------
    def decide_batch(self, token_ids: np.ndarray, lens: np.ndarray, thr: np.ndarray, want_cos: bool = True,
                     want_emb: bool = False, out: Optional[DecideResult] = None) -> DecideResult:
        token_ids = _as(token_ids, np.int32)
        b, n, t = token_ids.shape
        lens = _as(lens, np.int32).reshape(b, n)
        thr = _as(thr, np.float32).reshape(b)
        if out is None:
            out = DecideResult(np.empty(b, np.int32), np.empty(b, np.int32),
                               np.empty((b, n, n), np.float32) if want_cos else None,
                               np.empty((b, n, self.dim), np.float32) if want_emb else None)
        _check(self._lib.qc_decide_batch(self._h, _ptr(token_ids), _ptr(lens), b, n, t, _ptr(thr),
                                         _ptr(out.ref_ok), _ptr(out.medoid), _ptr(out.cos), _ptr(out.emb)))
        return out

    def decide_batch_submit(self, slot: int, token_ids: np.ndarray, lens: np.ndarray, thr: np.ndarray,
                            out: DecideResult) -> None:
        """Pipelined form: enqueue one batch on slot 0..3 (QC_PIPELINE_SLOTS); arrays must be C-contiguous int32/float32 and stay
        alive and untouched until ``decide_batch_wait(slot)`` returns (no copies are made here)."""
        b, n, t = token_ids.shape
        for a, dt in ((token_ids, np.int32), (lens, np.int32), (thr, np.float32)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("submit needs C-contiguous arrays of the exact dtype (they are read asynchronously)")
        _check(self._lib.qc_decide_batch_submit(self._h, int(slot), _ptr(token_ids), _ptr(lens), b, n, t, _ptr(thr),
                                                _ptr(out.ref_ok), _ptr(out.medoid), _ptr(out.cos)))

    def decide_ragged_submit(self, slot: int, tokens: np.ndarray, offsets: np.ndarray, n: int, thr: np.ndarray,
                             out: "DecideResult") -> None:
        """Pipelined CSR form; arrays must be C-contiguous and stay alive and untouched until ``decide_batch_wait``."""
        for a, dt in ((tokens, np.int32), (offsets, np.int32), (thr, np.float32), (out.ref_ok, np.int32), (out.medoid, np.int32)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("pipelined calls take C-contiguous arrays of the exact dtype (no copies are made)")
        b = (offsets.shape[0] - 1) // int(n)
        _check(self._lib.qc_decide_ragged_submit(self._h, int(slot), _ptr(tokens), _ptr(offsets), b, int(n), _ptr(thr),
                                                 _ptr(out.ref_ok), _ptr(out.medoid), _ptr(out.cos)))

    def decide_batch_wait(self, slot: int) -> None:
        _check(self._lib.qc_decide_batch_wait(self._h, int(slot)))

    def embed_batch(self, token_ids: np.ndarray, lens: np.ndarray) -> np.ndarray:
        token_ids = _as(token_ids, np.int32)
        r, t = token_ids.shape
        lens = _as(lens, np.int32).reshape(r)
        emb = np.empty((r, self.dim), np.float32)
        _check(self._lib.qc_embed_batch(self._h, _ptr(token_ids), _ptr(lens), r, t, _ptr(emb)))
        return emb

    def embed_ragged(self, tokens: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """CSR form of embed_batch: candidate r is tokens[offsets[r]:offsets[r+1]]; nothing is padded."""
        tokens = _as(tokens, np.int32).reshape(-1)
        offsets = _as(offsets, np.int32).reshape(-1)
        r = offsets.shape[0] - 1
        emb = np.empty((r, self.dim), np.float32)
        _check(self._lib.qc_embed_ragged(self._h, _ptr(tokens), _ptr(offsets), r, _ptr(emb)))
        return emb

    def decide_ragged(self, tokens: np.ndarray, offsets: np.ndarray, n: int, thr: np.ndarray, want_cos: bool = True,
                      want_emb: bool = False) -> "DecideResult":
        """CSR form of decide_batch: B = (len(offsets) - 1) / n decisions of n candidates each."""
        tokens = _as(tokens, np.int32).reshape(-1)
        offsets = _as(offsets, np.int32).reshape(-1)
        r = offsets.shape[0] - 1
        if r % n:
            raise ValueError("len(offsets) - 1 must be a multiple of the pool size")
        b = r // n
        thr = _as(thr, np.float32).reshape(b)
        res = DecideResult(np.empty(b, np.int32), np.empty(b, np.int32),
                           np.empty((b, n, n), np.float32) if want_cos else None,
                           np.empty((b, n, self.dim), np.float32) if want_emb else None)
        _check(self._lib.qc_decide_ragged(self._h, _ptr(tokens), _ptr(offsets), b, int(n), _ptr(thr), _ptr(res.ref_ok),
                                          _ptr(res.medoid), _ptr(res.cos) if want_cos else None,
                                          _ptr(res.emb) if want_emb else None))
        return res

    def similarity_batch(self, emb: np.ndarray, thr: np.ndarray, want_cos: bool = True):
        emb = _as(emb, np.float32)
        b, n, d = emb.shape
        thr = _as(thr, np.float32).reshape(b)
        ref_ok = np.empty(b, np.int32)
        medoid = np.empty(b, np.int32)
        cos = np.empty((b, n, n), np.float32) if want_cos else None
        _check(self._lib.qc_similarity_batch(self._h, _ptr(emb), b, n, d, _ptr(thr), _ptr(ref_ok), _ptr(medoid),
                                             _ptr(cos)))
        return ref_ok, medoid, cos

    def first_match_batch(self, query: np.ndarray, rows: np.ndarray, thr: np.ndarray, want_cos: bool = True):
        query = _as(query, np.float32)
        rows = _as(rows, np.float32)
        q, d = query.shape
        m = rows.shape[1]
        assert rows.shape == (q, m, d)
        thr = _as(thr, np.float32).reshape(q)
        idx = np.empty(q, np.int32)
        cos = np.empty((q, m), np.float32) if want_cos else None
        _check(self._lib.qc_first_match_batch(self._h, _ptr(query), _ptr(rows) if m else None, q, m, d, _ptr(thr),
                                              _ptr(idx), _ptr(cos) if m else None))
        return idx, cos

    def chunk_mean(self, chunks: np.ndarray) -> np.ndarray:
        chunks = _as(chunks, np.float32)
        g, c, d = chunks.shape
        out = np.empty((g, d), np.float32)
        _check(self._lib.qc_chunk_mean(self._h, _ptr(chunks), g, c, d, _ptr(out)))
        return out

    # -- device entry points (addresses of device buffers on this context's GPU) -------------
    def decide_batch_dev(self, d_token_ids, d_lens, b: int, n: int, t: int, d_thr, d_ref_ok=0, d_medoid=0,
                         d_cos=0, d_emb=0, stream=0) -> None:
        _check(self._lib.qc_decide_batch_dev(self._h, _dev(d_token_ids), _dev(d_lens), b, n, t, _dev(d_thr),
                                             _dev(d_ref_ok), _dev(d_medoid), _dev(d_cos), _dev(d_emb), _dev(stream)))

    def embed_batch_dev(self, d_token_ids, d_lens, r: int, t: int, d_emb, stream=0) -> None:
        _check(self._lib.qc_embed_batch_dev(self._h, _dev(d_token_ids), _dev(d_lens), r, t, _dev(d_emb), _dev(stream)))

    def similarity_batch_dev(self, d_emb, b: int, n: int, d: int, d_thr, d_ref_ok=0, d_medoid=0, d_cos=0,
                             stream=0) -> None:
        _check(self._lib.qc_similarity_batch_dev(self._h, _dev(d_emb), b, n, d, _dev(d_thr), _dev(d_ref_ok),
                                                 _dev(d_medoid), _dev(d_cos), _dev(stream)))

    def capture(self, stream, record) -> "Graph":
        """Record the `_dev` calls `record()` makes on `stream` (a non-default stream handle) into a CUDA graph."""
        _check(self._lib.qc_capture_begin(self._h, _dev(stream)))
        try:
            record()
        finally:
            h = ctypes.c_void_p()
            st = self._lib.qc_capture_end(self._h, _dev(stream), ctypes.byref(h))
        _check(st)
        return Graph(self._lib, h)

    def lesson_cache(self, capacity: int = 128, dim: Optional[int] = None) -> "LessonCache":
        """Device-resident rows of one agent's lesson list (qc_lesson_cache_*); see LessonCache."""
        return LessonCache(self, capacity, dim or self.dim)

    def token_split(self, world: int, rank: int, max_rows: int) -> "TokenSplit":
        """Exchange object for candidates sharded on T over `world` GPUs; see TokenSplit."""
        return TokenSplit(self, world, rank, max_rows)

    def split_pool(self, world: int, rank: int, max_decisions: int, pool_local: int) -> "SplitPool":
        """Gathered-buffer object for a pool split over `world` GPUs (one process each); see SplitPool."""
        return SplitPool(self, world, rank, max_decisions, pool_local)

    # -- introspection ---------------------------------------------------------------------------
    def check_status(self) -> None:
        """Synchronise and raise if an earlier asynchronous ``*_dev`` call flagged a device-side error."""
        _check(self._lib.qc_check_status(self._h))

    def last_hang(self) -> dict:
        info = (ctypes.c_int32 * 4)()
        _check(self._lib.qc_last_hang(self._h, info))
        return {"flag": info[0], "tag": info[1], "block": info[2], "thread": info[3]}

    def launch_count(self) -> int:
        return int(self._lib.qc_launch_count(self._h))

    def profile_enable(self, on: bool = True) -> None:
        _check(self._lib.qc_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self) -> None:
        _check(self._lib.qc_profile_reset(self._h))

    def kernel_stats(self) -> dict:
        ms = (ctypes.c_double * len(QC_K_NAMES))()
        nl = (ctypes.c_uint64 * len(QC_K_NAMES))()
        _check(self._lib.qc_kernel_stats(self._h, ms, nl))
        return {name: {"ms": float(ms[i]), "launches": int(nl[i])} for i, name in enumerate(QC_K_NAMES)}

    def encoder_info(self) -> dict:
        cg, tb, ws = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _check(self._lib.qc_encoder_info(self._h, ctypes.byref(cg), ctypes.byref(tb), ctypes.byref(ws)))
        return {"cta_group": cg.value, "token_block": tb.value, "w_stages": ws.value, "w_multicast_pairs": self.multicast()}

    STALL_FIELDS = ("mma_total", "mma_wait_w", "mma_wait_tokens", "mma_wait_acc", "epi_total", "epi_wait_acc",
                    "gather_total", "gather_wait_free", "wprod_total", "wprod_wait_free", "gather_wait_loads", "mma_issue", "mma_commit", "gather_fence", "aux14", "aux15")

    def stall_profile(self, enable: bool = True, read: bool = False):
        """K1 tuning aid: per-role barrier stall cycles of the last encoder launch (averaged over CTAs)."""
        out = (ctypes.c_double * 16)() if read else None
        _check(self._lib.qc_encoder_stall_profile(self._h, 1 if enable else 0, out))
        if out is None:
            return None
        return {name: float(out[i]) for i, name in enumerate(self.STALL_FIELDS)}

    def set_packing(self, mode: int) -> None:
        """Ragged-batch token packing on the host entry points: 0 auto, 1 never, 2 whenever possible."""
        _check(self._lib.qc_encoder_set_packing(self._h, int(mode)))

    def token_block(self, t: int) -> int:
        """Tokens per K1 block for sequences of `t` token slots (the tiling chosen for that T)."""
        tb = ctypes.c_int32()
        _check(self._lib.qc_encoder_token_block(self._h, int(t), ctypes.byref(tb)))
        return tb.value

    def hint_block_map(self, d_block_map, nslots: int) -> None:
        """Ids of the non-empty token blocks (device int32 array) for the NEXT encoder launch; see the header."""
        _check(self._lib.qc_encoder_hint_block_map(self._h, _dev(d_block_map), int(nslots)))

    def set_similarity_path(self, path: int) -> None:
        """K3 kernel for pools of 9..64: 0 auto, 1 mma.sync, 2 tcgen05 (test/tuning hook)."""
        _check(self._lib.qc_similarity_set_path(self._h, int(path)))

    def set_variant(self, token_block: int, cta_group: int, w_bk: int = 64, gather: int = 0) -> None:
        _check(self._lib.qc_encoder_set_variant(self._h, int(token_block), int(cta_group), int(w_bk), int(gather)))

    def set_multicast(self, pairs: int) -> None:
        """K1 W-tile multicast degree: 0 default, 1 off, 2 / 4 CTA pairs per cluster (see the header)."""
        _check(self._lib.qc_encoder_set_multicast(self._h, int(pairs)))

    def multicast(self) -> int:
        v = ctypes.c_int32()
        _check(self._lib.qc_encoder_multicast(self._h, ctypes.byref(v)))
        return v.value

    def set_cta_group(self, cta_group: int) -> None:
        _check(self._lib.qc_encoder_set_cta_group(self._h, int(cta_group)))


def _as(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


QC_SPLIT_HANDLE_BYTES = 64


class SplitPool:
    """A decision pool split across the GPUs of one node (qc_split_*).

    Each rank holds `pool_local` consecutive members of every pool.  ``handle()`` is this rank's 64-byte
    export; ``attach(rank, handle)`` maps a peer's.  ``decide_dev`` embeds the local candidates, stores the
    normalised rows into every rank's gathered buffer over peer memory, and decides on the full pool."""

    def __init__(self, ctx: "Context", world: int, rank: int, max_decisions: int, pool_local: int):
        self._ctx = ctx
        self._lib = ctx._lib
        self.world, self.rank, self.pool_local = int(world), int(rank), int(pool_local)
        h = ctypes.c_void_p()
        _check(self._lib.qc_split_create(ctx._h, self.world, self.rank, int(max_decisions), self.pool_local,
                                         ctypes.byref(h)))
        self._h = h

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(QC_SPLIT_HANDLE_BYTES)
        _check(self._lib.qc_split_export(self._h, buf))
        return bytes(buf.raw)

    def attach(self, peer_rank: int, handle: bytes) -> None:
        if len(handle) != QC_SPLIT_HANDLE_BYTES:
            raise ValueError("split-pool handle must be 64 bytes")
        _check(self._lib.qc_split_attach(self._h, int(peer_rank), handle))

    def decide_dev(self, d_token_ids, d_lens, b: int, t: int, d_thr, d_ref_ok=0, d_medoid=0, d_cos=0,
                   stream=0) -> None:
        _check(self._lib.qc_decide_split_dev(self._ctx._h, self._h, _dev(d_token_ids), _dev(d_lens), int(b), int(t),
                                             _dev(d_thr), _dev(d_ref_ok), _dev(d_medoid), _dev(d_cos), _dev(stream)))

    def gathered(self):
        """(device address, decisions) of the gathered rows [decisions, world * pool_local, dim] of the last call."""
        p, n = ctypes.c_void_p(), ctypes.c_int32()
        _check(self._lib.qc_split_gathered(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def close(self) -> None:
        if self._h:
            self._lib.qc_split_destroy(self._h)
            self._h = None


class TokenSplit(SplitPool):
    """Long candidates sharded on T across the GPUs of one node (qc_tsplit_*).

    Rank p passes the p-th token slice of every candidate; every rank ends up with the full embeddings."""

    def __init__(self, ctx: "Context", world: int, rank: int, max_rows: int):
        self._ctx = ctx
        self._lib = ctx._lib
        self.world, self.rank = int(world), int(rank)
        h = ctypes.c_void_p()
        _check(self._lib.qc_tsplit_create(ctx._h, self.world, self.rank, int(max_rows), ctypes.byref(h)))
        self._h = h

    def embed_dev(self, d_token_ids, d_lens, r: int, t_local: int, d_emb, stream=0) -> None:
        _check(self._lib.qc_embed_tsplit_dev(self._ctx._h, self._h, _dev(d_token_ids), _dev(d_lens), int(r),
                                             int(t_local), _dev(d_emb), _dev(stream)))

    def decide_dev(self, d_token_ids, d_lens, b: int, n: int, t_local: int, d_thr, d_ref_ok=0, d_medoid=0, d_cos=0,
                   d_emb=0, stream=0) -> None:
        _check(self._lib.qc_decide_tsplit_dev(self._ctx._h, self._h, _dev(d_token_ids), _dev(d_lens), int(b), int(n),
                                              int(t_local), _dev(d_thr), _dev(d_ref_ok), _dev(d_medoid), _dev(d_cos),
                                              _dev(d_emb), _dev(stream)))

    def gathered(self):
        raise NotImplementedError("token-split objects hold partial sums, not gathered rows")


class LessonCache:
    """One agent's lesson embeddings resident in HBM, in list order (qc_lesson_cache_*):
    a dedup call uploads one query row instead of the whole [M x D] list."""

    def __init__(self, ctx: "Context", capacity: int, dim: int):
        self._ctx = ctx
        self._lib = ctx._lib
        self.capacity, self.dim = int(capacity), int(dim)
        h = ctypes.c_void_p()
        _check(self._lib.qc_lesson_cache_create(ctx._h, self.capacity, self.dim, ctypes.byref(h)))
        self._h = h

    def __len__(self) -> int:
        return int(self._lib.qc_lesson_cache_count(self._h))

    def put(self, index: int, rows: np.ndarray) -> None:
        rows = _as(rows, np.float32).reshape(-1, self.dim)
        _check(self._lib.qc_lesson_cache_put(self._h, int(index), _ptr(rows), rows.shape[0]))

    def append(self, rows: np.ndarray) -> None:
        self.put(len(self), rows)

    def gather(self, src_index) -> None:
        idx = _as(np.asarray(src_index, dtype=np.int32).reshape(-1), np.int32)
        _check(self._lib.qc_lesson_cache_gather(self._h, _ptr(idx) if idx.size else None, idx.size))

    def match(self, query: np.ndarray, thr: float, want_cos: bool = False):
        """-> (index of the first resident row with cos >= thr or -1, cosines [count] or None)."""
        q = _as(query, np.float32).reshape(self.dim)
        idx = ctypes.c_int32(-1)
        cos = np.empty(len(self), np.float32) if want_cos else None
        _check(self._lib.qc_lesson_cache_match(self._h, _ptr(q), ctypes.c_float(thr), ctypes.byref(idx), _ptr(cos)))
        return idx.value, cos

    def close(self) -> None:
        if self._h:
            self._lib.qc_lesson_cache_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self._ctx._h:
                self.close()
        except Exception:
            pass


class Graph:
    """A recorded loop of `_dev` calls (qc_capture_*): ``launch(stream)`` replays it with one launch."""

    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle

    def launch(self, stream) -> None:
        _check(self._lib.qc_graph_launch(self._h, _dev(stream)))

    def close(self) -> None:
        if self._h:
            self._lib.qc_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
