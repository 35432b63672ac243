"""Python side of the mock Erlang runtime (see erl_nif.h / mock_erts.c in this directory): builds the UNMODIFIED NIF
shim quoracle_b200/nif/qconsensus_nif.c against the mock, loads it, and plays the VM -- packs argument terms, calls
`Elixir.Quoracle.Native` functions by name/arity, unpacks result terms.  TEST INFRASTRUCTURE only."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
NIF_SRC = ROOT / "quoracle_b200" / "nif" / "qconsensus_nif.c"
LIBDIR = ROOT / "quoracle_b200" / "lib"
OUT = HERE / "_build" / "qconsensus_nif_mock.so"

K_INT, K_U64, K_ATOM, K_TUPLE, K_BIN, K_RES, K_BADARG = 1, 2, 3, 4, 5, 6, 7


def build(force: bool = False) -> Path:
    deps = [NIF_SRC, HERE / "mock_erts.c", HERE / "erl_nif.h", ROOT / "include" / "qconsensus.h"]
    if not force and OUT.exists() and all(OUT.stat().st_mtime > d.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    tmp = OUT.with_suffix(f".{os.getpid()}.tmp.so")
    cmd = ["/usr/bin/gcc", "-O1", "-g", "-Wall", "-Wextra", "-shared", "-fPIC", "-pthread",
           "-I", str(HERE), "-I", str(ROOT / "include"), str(NIF_SRC), str(HERE / "mock_erts.c"),
           "-L", str(LIBDIR), "-lqconsensus", f"-Wl,-rpath,{LIBDIR}", "-o", str(tmp)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building the NIF against the mock runtime failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, OUT)
    return OUT


class Atom(str):
    def __repr__(self):
        return ":" + str(self)


class Resource:
    """A resource term kept alive in the VM's own long-lived environment."""

    def __init__(self, term):
        self.term = term


class BadArg(Exception):
    pass


class MockVM:
    _loaded = False

    def __init__(self):
        self.lib = ctypes.CDLL(str(build()))
        L, vp, sz = self.lib, ctypes.c_void_p, ctypes.c_size_t
        L.mock_env_new.restype = vp
        L.mock_env_free.argtypes = [vp]
        L.mock_make_binary.restype = vp
        L.mock_make_binary.argtypes = [vp, ctypes.c_char_p, sz]
        L.enif_make_int.restype = vp
        L.enif_make_int.argtypes = [vp, ctypes.c_int]
        L.enif_make_atom.restype = vp
        L.enif_make_atom.argtypes = [vp, ctypes.c_char_p]
        L.mock_copy_term.restype = vp
        L.mock_copy_term.argtypes = [vp, vp]
        L.mock_term_kind.argtypes = [vp]
        L.mock_tuple_arity.argtypes = [vp]
        L.mock_tuple_elem.restype = vp
        L.mock_tuple_elem.argtypes = [vp, ctypes.c_int]
        L.mock_atom_name.restype = ctypes.c_char_p
        L.mock_atom_name.argtypes = [vp]
        L.mock_binary_data.restype = vp
        L.mock_binary_data.argtypes = [vp]
        L.mock_binary_size.restype = sz
        L.mock_binary_size.argtypes = [vp]
        L.mock_get_uint64.restype = ctypes.c_ulonglong
        L.mock_get_uint64.argtypes = [vp]
        L.mock_live_resources.restype = ctypes.c_long
        L.mock_call.restype = vp
        L.mock_call.argtypes = [vp, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
        L.mock_module_name.restype = ctypes.c_char_p
        L.mock_func_name.restype = ctypes.c_char_p
        L.mock_func_name.argtypes = [ctypes.c_int]
        L.mock_func_arity.argtypes = [ctypes.c_int]
        L.mock_func_flags.argtypes = [ctypes.c_int]
        # on_load runs once per loaded library, as in a VM (a second run would re-open the resource types and orphan every
        # live resource); every MockVM of the process shares the one loaded library
        if not MockVM._loaded:
            if L.mock_load() != 0:
                raise RuntimeError("NIF on_load failed")
            MockVM._loaded = True
        self.heap = L.mock_env_new()          # where resource terms the test holds on to live

    # -- module table --------------------------------------------------------------------------------
    def module(self) -> str:
        return self.lib.mock_module_name().decode()

    def exports(self) -> dict:
        return {(self.lib.mock_func_name(i).decode(), self.lib.mock_func_arity(i)): self.lib.mock_func_flags(i)
                for i in range(self.lib.mock_num_funcs())}

    def live_resources(self) -> int:
        return int(self.lib.mock_live_resources())

    # -- terms ---------------------------------------------------------------------------------------
    def _pack(self, env, a):
        if isinstance(a, Resource):
            return self.lib.mock_copy_term(env, a.term)
        if isinstance(a, Atom):
            return self.lib.enif_make_atom(env, str(a).encode())
        if a is None:
            return self.lib.enif_make_atom(env, b"nil")
        if isinstance(a, bool):
            return self.lib.enif_make_atom(env, b"true" if a else b"false")
        if isinstance(a, int):
            return self.lib.enif_make_int(env, a)
        if isinstance(a, (bytes, bytearray, memoryview)):
            b = bytes(a)
            return self.lib.mock_make_binary(env, b, len(b))
        raise TypeError(f"cannot make a term from {type(a)}")

    def _unpack(self, t):
        k = self.lib.mock_term_kind(t)
        if k == K_ATOM:
            name = self.lib.mock_atom_name(t).decode()
            return None if name == "nil" else Atom(name)
        if k == K_U64:
            return int(self.lib.mock_get_uint64(t))
        if k == K_INT:
            v = int(self.lib.mock_get_uint64(t))               # two's complement of the signed value
            return v - (1 << 64) if v >= (1 << 63) else v
        if k == K_BIN:
            n = self.lib.mock_binary_size(t)
            return ctypes.string_at(self.lib.mock_binary_data(t), n) if n else b""
        if k == K_TUPLE:
            return tuple(self._unpack(self.lib.mock_tuple_elem(t, i)) for i in range(self.lib.mock_tuple_arity(t)))
        if k == K_RES:
            return Resource(self.lib.mock_copy_term(self.heap, t))
        if k == K_BADARG:
            raise BadArg()
        raise ValueError(f"unknown term kind {k}")

    def call(self, name: str, *args):
        """Quoracle.Native.name(args...) in a fresh process environment (freed afterwards, as a GC would)."""
        env = self.lib.mock_env_new()
        try:
            argv = (ctypes.c_void_p * max(1, len(args)))(*[self._pack(env, a) for a in args])
            t = self.lib.mock_call(env, name.encode(), len(args), argv)
            if not t:
                raise AttributeError(f"{self.module()}.{name}/{len(args)} is not exported by the NIF")
            return self._unpack(t)
        finally:
            self.lib.mock_env_free(env)

    def collect(self):
        """Drop every term the test still holds (a full GC): resources whose last reference that was are destroyed."""
        self.lib.mock_env_free(self.heap)
        self.heap = self.lib.mock_env_new()
