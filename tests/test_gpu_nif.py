"""The UNMODIFIED NIF shim executed on the B200 against the mock Erlang runtime (tests/mock_erts/): packed
little-endian binaries in, `{:ok, binaries}` / `{:error, atom}` out -- what the Elixir collector and the
`:embedding_fn` closure of elixir/quoracle_native.ex would see -- compared byte for byte with the ctypes path over
the same C ABI.  Seam being served: lib/quoracle/actions/consensus_rules.ex:302-327 (2-arity embedding_fn,
`{:ok, %{embedding: ...}}` | `{:error, atom}`)."""
import threading

import numpy as np
import pytest

from quoracle_b200 import native, synth
from tests import mock_erts
from tests.mock_erts import Atom, Resource

pytestmark = pytest.mark.gpu
OK, ERROR = Atom("ok"), Atom("error")
DIM, VOCAB = 768, 4096


@pytest.fixture(scope="module")
def world():
    vm = mock_erts.MockVM()
    w = synth.make_encoder(DIM, vocab=VOCAB, seed=synth.BASE_SEED + 11)
    tag, ctx = vm.call("ctx_create", 0, VOCAB, DIM, w.table_bf16.tobytes(), w.w_bf16.tobytes())
    assert tag == OK and isinstance(ctx, Resource)
    ref = native.Context(0, w.table_bf16, w.w_bf16)
    yield vm, ctx, ref
    ref.close()


def test_decide_batch_bytes_equal_the_ctypes_path(world):
    vm, ctx, ref = world
    batch = synth.make_decisions(24, 8, 200, vocab=VOCAB, seed=5, ragged=True, thresholds=(0.85, 0.995, 0.9995))
    want = ref.decide_batch(batch.token_ids, batch.lens, batch.thr, want_cos=True)
    tag, ok, med, cos = vm.call("decide_batch", ctx, batch.token_ids.tobytes(), batch.lens.tobytes(), 24, 8, 200,
                                batch.thr.tobytes(), 1)
    assert tag == OK
    assert ok == want.ref_ok.tobytes() and med == want.medoid.tobytes() and cos == want.cos.tobytes()
    tag, ok2, med2, cos2 = vm.call("decide_batch", ctx, batch.token_ids.tobytes(), batch.lens.tobytes(), 24, 8, 200,
                                   batch.thr.tobytes(), 0)
    assert (tag, ok2, med2, cos2) == (OK, ok, med, None)          # cos :: nil when not asked for
    assert len(set(np.frombuffer(ok, np.int32).tolist())) == 2     # both verdicts of the reference rule occur


def test_embed_ragged_and_first_match(world):
    vm, ctx, ref = world
    rng = np.random.default_rng(3)
    lens = np.array([5, 300, 0, 77, 1, 129], np.int32)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tokens = rng.integers(0, VOCAB, int(lens.sum())).astype(np.int32)
    want = ref.embed_ragged(tokens, offsets)
    tag, emb = vm.call("embed_ragged", ctx, tokens.tobytes(), offsets.tobytes())
    assert tag == OK and emb == want.tobytes()
    # lesson dedup on those rows: query = row 1, existing lessons = all rows
    q = want[1:2]
    rows = want[None, :, :]
    thr = np.array([0.999], np.float32)
    idx_want, _ = ref.first_match_batch(q, rows, thr)
    tag, idx = vm.call("first_match_batch", ctx, q.tobytes(), rows.tobytes(), 1, len(lens), DIM, thr.tobytes())
    assert tag == OK and idx == idx_want.tobytes() and np.frombuffer(idx, np.int32)[0] == 1
    # chunk mean and similarity through the same door
    tag, mean = vm.call("chunk_mean", ctx, want[[0, 1, 3]].tobytes(), 1, 3, DIM)
    assert tag == OK and mean == ref.chunk_mean(want[[0, 1, 3]][None]).tobytes()
    emb3 = np.ascontiguousarray(want[[0, 1, 3, 4, 5, 1]].reshape(2, 3, DIM))
    s_ok, s_med, s_cos = ref.similarity_batch(emb3, np.array([0.9, 0.9], np.float32))
    tag, ok, med, cos = vm.call("similarity_batch", ctx, emb3.tobytes(), 2, 3, DIM, np.array([0.9, 0.9], np.float32).tobytes())
    assert (tag, ok, med, cos) == (OK, s_ok.tobytes(), s_med.tobytes(), s_cos.tobytes())


def test_errors_are_atoms_and_the_context_survives(world):
    vm, ctx, ref = world
    tok = np.zeros((1, 2, 8), np.int32)
    tok[0, 1, 3] = VOCAB                                            # token id out of range
    lens = np.full((1, 2), 8, np.int32)
    thr = np.array([0.9], np.float32)
    assert vm.call("decide_batch", ctx, tok.tobytes(), lens.tobytes(), 1, 2, 8, thr.tobytes(), 0) == (ERROR, Atom("invalid_input"))
    assert vm.call("decide_batch", ctx, tok.tobytes()[:-4], lens.tobytes(), 1, 2, 8, thr.tobytes(), 0) == (ERROR, Atom("invalid_input"))
    long = np.zeros((1, 1, 8193), np.int32)
    assert vm.call("embed_batch", ctx, long.tobytes(), np.array([8193], np.int32).tobytes(), 1, 8193) == \
        (ERROR, Atom("context_length_exceeded"))
    assert vm.call("similarity_batch", ctx, np.zeros((1, 65, 4), np.float32).tobytes(), 1, 65, 4, thr.tobytes()) == \
        (ERROR, Atom("bad_request"))
    tok[0, 1, 3] = 7
    tag, *_ = vm.call("decide_batch", ctx, tok.tobytes(), lens.tobytes(), 1, 2, 8, thr.tobytes(), 0)
    assert tag == OK                                                 # usable after every error above


def test_pipelined_submit_wait_through_the_nif(world):
    """Four batches in flight (the e2e path of bench.py), argument binaries freed right after submit: results equal
    the synchronous NIF call."""
    vm, ctx, ref = world
    batches = [synth.make_decisions(16, 8, 256, vocab=VOCAB, seed=70 + i, ragged=(i % 2 == 0)) for i in range(6)]
    want = [vm.call("decide_batch", ctx, x.token_ids.tobytes(), x.lens.tobytes(), 16, 8, 256, x.thr.tobytes(), 1)
            for x in batches]
    got = [None] * len(batches)
    for i, x in enumerate(batches):
        k = i % 4
        if i >= 4:
            got[i - 4] = vm.call("decide_batch_wait", ctx, k)
        assert vm.call("decide_batch_submit", ctx, k, x.token_ids.tobytes(), x.lens.tobytes(), 16, 8, 256,
                       x.thr.tobytes(), 1) == OK
    for i in range(len(batches) - 4, len(batches)):
        got[i] = vm.call("decide_batch_wait", ctx, i % 4)
    assert got == want
    assert vm.call("decide_batch_wait", ctx, 0) == (ERROR, Atom("invalid_input"))       # nothing in flight
    # CSR form
    x = batches[0]
    lens = x.lens.reshape(-1)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tokens = np.concatenate([x.token_ids.reshape(-1, 256)[r, :lens[r]] for r in range(lens.size)]).astype(np.int32)
    sync = vm.call("decide_ragged", ctx, tokens.tobytes(), offsets.tobytes(), 8, x.thr.tobytes(), 1)
    assert vm.call("decide_ragged_submit", ctx, 2, tokens.tobytes(), offsets.tobytes(), 8, x.thr.tobytes(), 1) == OK
    assert vm.call("decide_ragged_submit", ctx, 2, tokens.tobytes(), offsets.tobytes(), 8, x.thr.tobytes(), 1) == \
        (ERROR, Atom("invalid_input"))                                                  # slot busy
    assert vm.call("decide_batch_wait", ctx, 2) == sync


def test_concurrent_callers_share_one_context(world):
    """Eight 'agent processes' (threads = dirty schedulers) call embed_ragged on the SAME ctx at once, as every agent's
    :embedding_fn closure does: the per-resource mutex serialises them, every caller gets its own correct rows."""
    vm, ctx, ref = world
    rng = np.random.default_rng(17)
    jobs = []
    for a in range(8):
        lens = rng.integers(1, 400, 5).astype(np.int32)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        tokens = rng.integers(0, VOCAB, int(lens.sum())).astype(np.int32)
        jobs.append((tokens, offsets, ref.embed_ragged(tokens, offsets).tobytes()))
    errors = []

    def agent(tokens, offsets, want):
        for _ in range(20):
            r = vm.call("embed_ragged", ctx, tokens.tobytes(), offsets.tobytes())
            if r != (OK, want):
                errors.append(r[0])

    threads = [threading.Thread(target=agent, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors


def test_resource_lifetime():
    """The context lives exactly as long as a term refers to it; an in-flight slot is drained by the destructor."""
    vm = mock_erts.MockVM()
    base = vm.live_resources()
    w = synth.make_encoder(384, vocab=1024, seed=1)
    tag, ctx = vm.call("ctx_create", 0, 1024, 384, w.table_bf16.tobytes(), w.w_bf16.tobytes())
    assert tag == OK and vm.live_resources() == base + 1
    x = synth.make_decisions(4, 3, 64, vocab=1024, seed=2)
    assert vm.call("decide_batch_submit", ctx, 1, x.token_ids.tobytes(), x.lens.tobytes(), 4, 3, 64, x.thr.tobytes(), 0) == OK
    del ctx
    vm.collect()                                                     # last reference gone with a batch still in flight
    assert vm.live_resources() == base


def test_lesson_cache_through_the_nif(world):
    """Rows resident on the device: append, match (first >= thr), replace a row, prune / reorder with gather."""
    vm, ctx, ref = world
    rng = np.random.default_rng(23)
    rows = rng.standard_normal((12, DIM)).astype(np.float32)
    tag, cache = vm.call("lesson_cache_create", ctx, 32, DIM)
    assert tag == OK
    thr = np.array([0.9], np.float32).tobytes()
    assert vm.call("lesson_cache_match", cache, rows[0].tobytes(), thr, 1) == (OK, -1, None)      # empty cache
    assert vm.call("lesson_cache_put", cache, 0, rows[:10].tobytes()) == OK
    assert vm.call("lesson_cache_put", cache, 11, rows[10:11].tobytes()) == (ERROR, Atom("invalid_input"))   # hole
    q = (3.0 * rows[7] + 0.01 * rows[2]).astype(np.float32)
    tag, idx, cos = vm.call("lesson_cache_match", cache, q.tobytes(), thr, 1)
    want_idx, want_cos = ref.first_match_batch(q[None], rows[None, :10], np.array([0.9], np.float32))
    assert (tag, idx) == (OK, 7) and idx == int(want_idx[0]) and cos == want_cos.tobytes()
    assert vm.call("lesson_cache_put", cache, 7, rows[11:12].tobytes()) == OK                       # merged: row replaced
    assert vm.call("lesson_cache_match", cache, q.tobytes(), thr, 0) == (OK, -1, None)
    order = np.array([9, 7, 0], np.int32)                                                           # prune + reorder
    assert vm.call("lesson_cache_gather", cache, order.tobytes()) == OK
    tag, idx, cos = vm.call("lesson_cache_match", cache, rows[11].tobytes(), thr, 1)
    assert (tag, idx) == (OK, 1) and len(cos) == 12
    assert vm.call("lesson_cache_gather", cache, np.array([5], np.int32).tobytes()) == (ERROR, Atom("invalid_input"))
    assert vm.call("lesson_cache_put", cache, 0, rows[0, :5].tobytes()) == (ERROR, Atom("invalid_input"))
