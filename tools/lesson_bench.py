"""Lesson dedup calls/s: one new lesson against M = 100 existing ones, D = 768.

  uncached   qc_first_match_batch: the [1 x M x D] rows travel host -> device on every call (round 1)
  cached     qc_lesson_cache_match: rows resident in HBM, one query row (D floats) uploaded per call
Both return the same index and cosines.  Wall clock over many calls (each call is synchronous, as the NIF's is)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quoracle_b200 import native  # noqa: E402


def main():
    m, d, calls = 100, 768, 2000
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((m, d)).astype(np.float32)
    qs = rng.standard_normal((64, d)).astype(np.float32)
    qs[::2] = rows[rng.integers(0, m, 32)] * 2.0                      # half of the queries match something
    thr = np.array([0.9], np.float32)
    ctx = native.Context(0)
    cache = ctx.lesson_cache(128, d)
    cache.append(rows)
    for k in range(64):                                               # same answers
        a, ac = ctx.first_match_batch(qs[k:k + 1], rows[None], thr)
        b, bc = cache.match(qs[k], 0.9, want_cos=True)
        assert int(a[0]) == b and np.array_equal(ac[0], bc)
    out = {}
    for name, fn in (("uncached_first_match_batch", lambda k: ctx.first_match_batch(qs[k % 64][None], rows[None], thr, want_cos=False)),
                     ("cached_lesson_cache_match", lambda k: cache.match(qs[k % 64], 0.9))):
        for k in range(50):
            fn(k)
        t0 = time.perf_counter()
        for k in range(calls):
            fn(k)
        dt = time.perf_counter() - t0
        out[name] = {"calls_per_s": calls / dt, "us_per_call": 1e6 * dt / calls}
    out["h2d_bytes_per_call"] = {"uncached": (m + 1) * d * 4 + 4, "cached": d * 4 + 4}
    out["speedup"] = out["cached_lesson_cache_match"]["calls_per_s"] / out["uncached_first_match_batch"]["calls_per_s"]
    out["shape"] = {"existing_lessons": m, "dim": d, "calls": calls}
    print(json.dumps(out))
    cache.close()
    ctx.close()


if __name__ == "__main__":
    main()
