"""The C ABI driven from plain C (examples/c_abi_decide.c, compiled with gcc against the in-tree library) must give
what the ctypes binding gives on the same inputs: the boundary is the C header, not the Python wrapper."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from quoracle_b200 import native

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def lcg_stream(n, seed=12345):
    """x_{k+1} = 1664525 x_k + 1013904223 (mod 2^32), vectorised: x_k = a^k x_0 + c (a^k - 1)/(a - 1)."""
    a = np.full(n, 1664525, np.uint32)
    ak = np.cumprod(a, dtype=np.uint32)                                   # a^1 .. a^n  (wraps mod 2^32)
    geo = np.cumsum(np.concatenate([[np.uint32(1)], ak[:-1]]), dtype=np.uint32)   # 1 + a + ... + a^(k-1)
    return ak * np.uint32(seed) + np.uint32(1013904223) * geo


def bf16_bits(x):
    return (x.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def test_c_program_matches_the_ctypes_path(tmp_path):
    exe = tmp_path / "c_abi_decide"
    lib = ROOT / "quoracle_b200" / "lib"
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "c_abi_decide.c"),
                    f"-L{lib}", "-lqconsensus", f"-Wl,-rpath,{lib}", "-o", str(exe)], check=True)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    got = json.loads(proc.stdout.strip().splitlines()[-1])

    v, d, b, n, t = 4096, 768, 6, 8, 200
    total = v * d + d * d + b * n * (t + 1)
    s = lcg_stream(total)
    to_unit = lambda u: ((u >> np.uint32(8)).astype(np.int64) - (1 << 23)).astype(np.float32) / np.float32(1 << 23)  # noqa: E731
    table = bf16_bits(to_unit(s[:v * d])).reshape(v, d)
    w = bf16_bits(to_unit(s[v * d:v * d + d * d]) / np.float32(27.7)).reshape(d, d)
    cur = v * d + d * d
    tok = np.full((b * n, t), -1, np.int32)
    lens = np.zeros(b * n, np.int32)
    for r in range(b * n):
        lens[r] = int(s[cur] % np.uint32(t + 1))
        cur += 1
        tok[r, :lens[r]] = (s[cur:cur + lens[r]] % np.uint32(v)).astype(np.int32)
        cur += lens[r]
    thr = np.full(b, 0.9995, np.float32)
    with native.Context(0, table, w) as ctx:
        res = ctx.decide_batch(tok.reshape(b, n, t), lens.reshape(b, n), thr, want_cos=True)
    assert got["abi"] >= 1 and got["dim"] == d
    assert got["medoid"] == res.medoid.tolist() and got["ref_ok"] == res.ref_ok.tolist()
    assert got["medoid_csr"] == res.medoid.tolist()
    assert abs(got["cos_0_0_1"] - float(res.cos[0, 0, 1])) < 1e-7
    assert got["max_cos_diff_padded_vs_csr"] < 1e-5
