"""First-contact GPU triage for K1 (not a test): prints per-variant error maps so a descriptor or
barrier mistake can be localised from one gpurun call."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np

from oracle import encoder as oenc
from quoracle_b200 import native, synth


def run(dim, cg, r, t, vocab=4096):
    w = synth.make_encoder(dim, vocab=vocab, seed=synth.BASE_SEED + dim)
    rng = np.random.default_rng(1)
    lens = np.full(r, t, np.int32)
    tok = rng.integers(0, vocab, (r, t)).astype(np.int32)
    want = oenc.encode(w.table_bf16, w.w_bf16, tok, lens)
    with native.Context(0, w.table_bf16, w.w_bf16) as ctx:
        ctx.set_cta_group(cg)
        t0 = time.time()
        try:
            got = ctx.embed_batch(tok, lens)
        except native.QcError as e:
            print(f"dim={dim} cg={cg} r={r} t={t}: ERROR {e}  hang={ctx.last_hang()}", flush=True)
            return
        dt = time.time() - t0
        err = np.abs(got - want)
        print(f"dim={dim} cg={cg} r={r} t={t}: max_err={err.max():.3e} mean_err={err.mean():.3e} "
              f"nan={np.isnan(got).sum()} time={dt:.3f}s info={ctx.encoder_info()}", flush=True)
        if err.max() > 1e-3:
            bad_rows = np.where(err.max(axis=1) > 1e-3)[0]
            print("   bad rows:", bad_rows[:20], "of", r)
            row = bad_rows[0]
            e = err[row]
            print("   row", row, "err by 64-feature group:", np.round(e.reshape(-1, 64).max(axis=1), 4))
            print("   got[:8]", got[row, :8], "want[:8]", want[row, :8])
            print("   cos(got,want) =", float(np.dot(got[row], want[row]) /
                                                (np.linalg.norm(got[row]) * np.linalg.norm(want[row]) + 1e-30)))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cases = [(384, 1), (768, 1), (1536, 1)] if which in ("all", "cg1") else []
    cases += [(768, 2), (1536, 2)] if which in ("all", "cg2") else []
    for dim, cg in cases:
        run(dim, cg, 3, 128 if dim != 1536 else 64)     # one block per candidate
        run(dim, cg, 5, 300)                              # several blocks, ragged tail block
        run(dim, cg, 400, 256)                            # multi-wave
