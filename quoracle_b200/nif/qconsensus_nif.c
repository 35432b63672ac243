/*
 * qconsensus_nif.c -- dirty-scheduler Erlang NIF over the C ABI in include/qconsensus.h.
 *
 * Module: Elixir.Quoracle.Native (elixir/quoracle_native.ex).  Every NIF is flagged
 * ERL_NIF_DIRTY_JOB_CPU_BOUND: a call blocks a dirty CPU scheduler thread for the duration of one batched
 * GPU step (~1 ms for 256 decisions), never a normal scheduler.  Arguments arrive as binaries packed by
 * the Elixir side (little-endian int32 / float32, row-major); results go back as {:ok, binary, ...} or
 * {:error, atom} with the atom qc_status_atom() maps the status to -- the atoms the reference's embedding
 * path already uses (lib/quoracle/models/embeddings.ex:49-50,184,228-229,300-331,344), so a failed call
 * makes ConsensusRules fall back to the mode exactly as a failed HTTP embedding does
 * (lib/quoracle/actions/consensus_rules.ex:294-297, lib/quoracle/consensus/result.ex:352-358).
 * Nothing here raises; badarg is returned only for malformed terms.
 *
 * Concurrency: include/qconsensus.h requires the calls on one context to be serialised, while the BEAM runs NIFs
 * of one resource from any number of dirty-scheduler threads at once (every agent process holds the same ctx in
 * its :embedding_fn closure, and the collector GenServer shares it).  Each context resource therefore carries an
 * ErlNifMutex that is held around every qc_* call on that context: two agents calling embed at the same time queue
 * behind each other (1 ms GPU steps) instead of racing on the context's staging buffers and plan.  The pipelined
 * pair decide_batch_submit / decide_batch_wait takes the lock only for the duration of each of its two calls, so
 * up to QC_PIPELINE_SLOTS batches are still in flight between them.
 *
 * NO BEAM IN THIS IMAGE: the shim cannot be loaded by a real VM here.  It IS executed, unmodified, against the
 * functional mock runtime in tests/mock_erts/ (erl_nif.h + mock_erts.c) by tests/test_gpu_nif.py on the B200:
 * packed binaries in, {:ok, binaries} / {:error, atom} out, resource lifetime and the per-context lock included.
 * Build the production NIF against OTP's own erl_nif.h ("Building the NIF").
 */
#include <erl_nif.h>
#include <string.h>

#include "qconsensus.h"

static ErlNifResourceType* QC_CTX_TYPE = NULL;

/* One in-flight batch of the pipelined pair: page-locked result buffers owned by the resource (the library copies
 * into them asynchronously; binaries are made from them at wait time) and the shape the wait has to return. */
typedef struct {
  int busy, b, n, want_cos;
  int32_t* ok;
  int32_t* med;
  float* cos;
  size_t ok_cap, cos_cap;       /* bytes */
} slot_res;

typedef struct {
  qc_ctx* ctx;
  ErlNifMutex* lock;            /* serialises every qc_* call on ctx (see the header comment) */
  slot_res slots[QC_PIPELINE_SLOTS];
} ctx_res;

static void ctx_dtor(ErlNifEnv* env, void* obj) {
  (void)env;
  ctx_res* r = (ctx_res*)obj;
  if (r->ctx) {
    for (int k = 0; k < QC_PIPELINE_SLOTS; ++k)
      if (r->slots[k].busy) qc_decide_batch_wait(r->ctx, k);      /* drain before the buffers go away */
    qc_ctx_destroy(r->ctx);
  }
  for (int k = 0; k < QC_PIPELINE_SLOTS; ++k) {
    qc_host_free(r->slots[k].ok);
    qc_host_free(r->slots[k].med);
    qc_host_free(r->slots[k].cos);
  }
  if (r->lock) enif_mutex_destroy(r->lock);
  r->ctx = NULL;
  r->lock = NULL;
}

static ERL_NIF_TERM err(ErlNifEnv* env, int status) {
  return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_atom(env, qc_status_atom(status)));
}

static int get_res(ErlNifEnv* env, ERL_NIF_TERM t, ctx_res** out) {
  ctx_res* r;
  if (!enif_get_resource(env, t, QC_CTX_TYPE, (void**)&r) || r->ctx == NULL) return 0;
  *out = r;
  return 1;
}

/* Argument check + handle; the caller brackets its qc_* call with LOCK(ctx) / UNLOCK(ctx). */
static int get_ctx(ErlNifEnv* env, ERL_NIF_TERM t, qc_ctx** out) {
  ctx_res* r;
  if (!get_res(env, t, &r)) return 0;
  *out = r->ctx;
  return 1;
}

/* The resource is recovered from the term again rather than threaded through every signature: one pointer chase. */
#define LOCK(TERM)                                            \
  ctx_res* lk_;                                               \
  get_res(env, (TERM), &lk_);                                 \
  enif_mutex_lock(lk_->lock)
#define UNLOCK() enif_mutex_unlock(lk_->lock)

static int is_nil(ErlNifEnv* env, ERL_NIF_TERM t) { return enif_is_identical(t, enif_make_atom(env, "nil")); }

/* ctx_create(device_id, vocab, dim, table_bf16 :: binary | nil, w_bf16 :: binary | nil) */
static ERL_NIF_TERM nif_ctx_create(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int device, vocab, dim;
  ErlNifBinary table, w;
  qc_encoder_cfg cfg;
  const qc_encoder_cfg* pcfg = NULL;
  if (argc != 5 || !enif_get_int(env, argv[0], &device) || !enif_get_int(env, argv[1], &vocab) ||
      !enif_get_int(env, argv[2], &dim))
    return enif_make_badarg(env);
  if (!is_nil(env, argv[3])) {
    if (!enif_inspect_binary(env, argv[3], &table) || !enif_inspect_binary(env, argv[4], &w)) return enif_make_badarg(env);
    if (table.size != (size_t)vocab * dim * 2 || w.size != (size_t)dim * dim * 2) return err(env, QC_ERR_INVALID_INPUT);
    cfg.vocab = vocab;
    cfg.dim = dim;
    cfg.table_bf16 = (const uint16_t*)table.data;
    cfg.w_bf16 = (const uint16_t*)w.data;
    pcfg = &cfg;
  }
  qc_ctx* ctx = NULL;
  int st = qc_ctx_create(&ctx, device, pcfg);
  if (st != QC_OK) return err(env, st);
  ctx_res* r = (ctx_res*)enif_alloc_resource(QC_CTX_TYPE, sizeof(ctx_res));
  memset(r, 0, sizeof(*r));
  r->ctx = ctx;
  r->lock = enif_mutex_create((char*)"qc_ctx");
  ERL_NIF_TERM term = enif_make_resource(env, r);
  enif_release_resource(r);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), term);
}

/* decide_batch(ctx, token_ids, lens, b, n, t, thr, want_cos :: 0 | 1)
 *   -> {:ok, ref_ok :: binary, medoid :: binary, cos :: binary | nil} | {:error, atom} */
static ERL_NIF_TERM nif_decide_batch(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary tok, lens, thr;
  int b, n, t, want_cos;
  if (argc != 8 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &tok) ||
      !enif_inspect_binary(env, argv[2], &lens) || !enif_get_int(env, argv[3], &b) || !enif_get_int(env, argv[4], &n) ||
      !enif_get_int(env, argv[5], &t) || !enif_inspect_binary(env, argv[6], &thr) || !enif_get_int(env, argv[7], &want_cos))
    return enif_make_badarg(env);
  if (b < 0 || n < 1 || t < 1 || tok.size != (size_t)b * n * t * 4 || lens.size != (size_t)b * n * 4 ||
      thr.size != (size_t)b * 4)
    return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM ok_t, med_t, cos_t = enif_make_atom(env, "nil");
  int32_t* ref_ok = (int32_t*)enif_make_new_binary(env, (size_t)b * 4, &ok_t);
  int32_t* medoid = (int32_t*)enif_make_new_binary(env, (size_t)b * 4, &med_t);
  float* cos = want_cos ? (float*)enif_make_new_binary(env, (size_t)b * n * n * 4, &cos_t) : NULL;
  LOCK(argv[0]);
  int st = qc_decide_batch(ctx, (const int32_t*)tok.data, (const int32_t*)lens.data, b, n, t, (const float*)thr.data,
                           ref_ok, medoid, cos, NULL);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple4(env, enif_make_atom(env, "ok"), ok_t, med_t, cos_t);
}

/* embed_batch(ctx, token_ids, lens, r, t) -> {:ok, emb :: binary (r*dim f32)} | {:error, atom} */
static ERL_NIF_TERM nif_embed_batch(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary tok, lens;
  int r, t;
  if (argc != 5 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &tok) ||
      !enif_inspect_binary(env, argv[2], &lens) || !enif_get_int(env, argv[3], &r) || !enif_get_int(env, argv[4], &t))
    return enif_make_badarg(env);
  const int dim = qc_ctx_dim(ctx);
  if (dim <= 0) return err(env, QC_ERR_NOT_CONFIGURED);
  if (r < 0 || t < 1 || tok.size != (size_t)r * t * 4 || lens.size != (size_t)r * 4) return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM emb_t;
  float* emb = (float*)enif_make_new_binary(env, (size_t)r * dim * 4, &emb_t);
  LOCK(argv[0]);
  int st = qc_embed_batch(ctx, (const int32_t*)tok.data, (const int32_t*)lens.data, r, t, emb);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), emb_t);
}

/* embed_ragged(ctx, tokens :: binary (int32), offsets :: binary (int32, r + 1 entries))
 *   -> {:ok, emb :: binary (r*dim f32)} | {:error, atom}
 * The collector's natural form: values of different lengths, nothing padded (qc_embed_ragged). */
static ERL_NIF_TERM nif_embed_ragged(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary tok, off;
  if (argc != 3 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &tok) ||
      !enif_inspect_binary(env, argv[2], &off))
    return enif_make_badarg(env);
  const int dim = qc_ctx_dim(ctx);
  if (dim <= 0) return err(env, QC_ERR_NOT_CONFIGURED);
  if (off.size < 4 || off.size % 4 != 0 || tok.size % 4 != 0) return err(env, QC_ERR_INVALID_INPUT);
  const int r = (int)(off.size / 4) - 1;
  const int32_t* offsets = (const int32_t*)off.data;
  if (offsets[r] < 0 || (size_t)offsets[r] * 4 > tok.size) return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM emb_t;
  float* emb = (float*)enif_make_new_binary(env, (size_t)r * dim * 4, &emb_t);
  LOCK(argv[0]);
  int st = qc_embed_ragged(ctx, (const int32_t*)tok.data, offsets, r, emb);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), emb_t);
}

/* decide_ragged(ctx, tokens, offsets, n, thr, want_cos :: 0 | 1)
 *   -> {:ok, ref_ok :: binary, medoid :: binary, cos :: binary | nil} | {:error, atom} */
static ERL_NIF_TERM nif_decide_ragged(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary tok, off, thr;
  int n, want_cos;
  if (argc != 6 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &tok) ||
      !enif_inspect_binary(env, argv[2], &off) || !enif_get_int(env, argv[3], &n) ||
      !enif_inspect_binary(env, argv[4], &thr) || !enif_get_int(env, argv[5], &want_cos))
    return enif_make_badarg(env);
  if (n < 1 || off.size < 4 || off.size % 4 != 0 || tok.size % 4 != 0) return err(env, QC_ERR_INVALID_INPUT);
  const int r = (int)(off.size / 4) - 1;
  if (r % n != 0) return err(env, QC_ERR_INVALID_INPUT);
  const int b = r / n;
  const int32_t* offsets = (const int32_t*)off.data;
  if (thr.size != (size_t)b * 4 || offsets[r] < 0 || (size_t)offsets[r] * 4 > tok.size) return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM ok_t, med_t, cos_t = enif_make_atom(env, "nil");
  int32_t* ref_ok = (int32_t*)enif_make_new_binary(env, (size_t)b * 4, &ok_t);
  int32_t* medoid = (int32_t*)enif_make_new_binary(env, (size_t)b * 4, &med_t);
  float* cos = want_cos ? (float*)enif_make_new_binary(env, (size_t)b * n * n * 4, &cos_t) : NULL;
  LOCK(argv[0]);
  int st = qc_decide_ragged(ctx, (const int32_t*)tok.data, offsets, b, n, (const float*)thr.data, ref_ok, medoid, cos, NULL);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple4(env, enif_make_atom(env, "ok"), ok_t, med_t, cos_t);
}

/* similarity_batch(ctx, emb, b, n, d, thr) -> {:ok, ref_ok, medoid, cos} | {:error, atom} */
static ERL_NIF_TERM nif_similarity_batch(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary emb, thr;
  int b, n, d;
  if (argc != 6 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &emb) ||
      !enif_get_int(env, argv[2], &b) || !enif_get_int(env, argv[3], &n) || !enif_get_int(env, argv[4], &d) ||
      !enif_inspect_binary(env, argv[5], &thr))
    return enif_make_badarg(env);
  if (b < 0 || n < 1 || d < 1 || emb.size != (size_t)b * n * d * 4 || thr.size != (size_t)b * 4)
    return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM ok_t, med_t, cos_t;
  int32_t* ref_ok = (int32_t*)enif_make_new_binary(env, (size_t)b * 4, &ok_t);
  int32_t* medoid = (int32_t*)enif_make_new_binary(env, (size_t)b * 4, &med_t);
  float* cos = (float*)enif_make_new_binary(env, (size_t)b * n * n * 4, &cos_t);
  LOCK(argv[0]);
  int st = qc_similarity_batch(ctx, (const float*)emb.data, b, n, d, (const float*)thr.data, ref_ok, medoid, cos);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple4(env, enif_make_atom(env, "ok"), ok_t, med_t, cos_t);
}

/* first_match_batch(ctx, query, rows, q, m, d, thr) -> {:ok, index :: binary (q int32)} | {:error, atom} */
static ERL_NIF_TERM nif_first_match_batch(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary query, rows, thr;
  int q, m, d;
  if (argc != 7 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &query) ||
      !enif_inspect_binary(env, argv[2], &rows) || !enif_get_int(env, argv[3], &q) || !enif_get_int(env, argv[4], &m) ||
      !enif_get_int(env, argv[5], &d) || !enif_inspect_binary(env, argv[6], &thr))
    return enif_make_badarg(env);
  if (q < 0 || m < 0 || d < 1 || query.size != (size_t)q * d * 4 || rows.size != (size_t)q * m * d * 4 ||
      thr.size != (size_t)q * 4)
    return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM idx_t;
  int32_t* idx = (int32_t*)enif_make_new_binary(env, (size_t)q * 4, &idx_t);
  LOCK(argv[0]);
  int st = qc_first_match_batch(ctx, (const float*)query.data, (const float*)rows.data, q, m, d,
                                (const float*)thr.data, idx, NULL);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), idx_t);
}

/* chunk_mean(ctx, chunks, g, c, d) -> {:ok, mean :: binary (g*d f32)} | {:error, atom} */
static ERL_NIF_TERM nif_chunk_mean(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  ErlNifBinary chunks;
  int g, c, d;
  if (argc != 5 || !get_ctx(env, argv[0], &ctx) || !enif_inspect_binary(env, argv[1], &chunks) ||
      !enif_get_int(env, argv[2], &g) || !enif_get_int(env, argv[3], &c) || !enif_get_int(env, argv[4], &d))
    return enif_make_badarg(env);
  if (g < 0 || c < 1 || d < 1 || chunks.size != (size_t)g * c * d * 4) return err(env, QC_ERR_INVALID_INPUT);
  ERL_NIF_TERM out_t;
  float* out = (float*)enif_make_new_binary(env, (size_t)g * d * 4, &out_t);
  LOCK(argv[0]);
  int st = qc_chunk_mean(ctx, (const float*)chunks.data, g, c, d, out);
  UNLOCK();
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), out_t);
}

/* ---- pipelined pair (qc_decide_batch_submit / qc_decide_ragged_submit / qc_decide_batch_wait) --------------------
 * submit copies nothing the VM owns past its return: the library stages non-pinned inputs into the slot's own
 * pinned area before it returns, and results land in the resource's page-locked buffers, from which `wait` makes
 * the result binaries.  So the argument binaries may be garbage-collected as soon as submit returns, and a slot
 * whose caller died is drained by the resource destructor. */
static int slot_prepare(slot_res* sl, int b, int n, int want_cos) {
  const size_t okb = (size_t)b * 4, cosb = want_cos ? (size_t)b * n * n * 4 : 0;
  if (okb > sl->ok_cap) {
    qc_host_free(sl->ok);
    qc_host_free(sl->med);
    sl->ok = (int32_t*)qc_host_alloc(okb);
    sl->med = (int32_t*)qc_host_alloc(okb);
    sl->ok_cap = (sl->ok && sl->med) ? okb : 0;
    if (!sl->ok_cap) return 0;
  }
  if (cosb > sl->cos_cap) {
    qc_host_free(sl->cos);
    sl->cos = (float*)qc_host_alloc(cosb);
    sl->cos_cap = sl->cos ? cosb : 0;
    if (!sl->cos_cap) return 0;
  }
  sl->b = b;
  sl->n = n;
  sl->want_cos = want_cos;
  return 1;
}

/* decide_batch_submit(ctx, slot, token_ids, lens, b, n, t, thr, want_cos) -> :ok | {:error, atom} */
static ERL_NIF_TERM nif_decide_batch_submit(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary tok, lens, thr;
  int slot, b, n, t, want_cos;
  if (argc != 9 || !get_res(env, argv[0], &r) || !enif_get_int(env, argv[1], &slot) ||
      !enif_inspect_binary(env, argv[2], &tok) || !enif_inspect_binary(env, argv[3], &lens) ||
      !enif_get_int(env, argv[4], &b) || !enif_get_int(env, argv[5], &n) || !enif_get_int(env, argv[6], &t) ||
      !enif_inspect_binary(env, argv[7], &thr) || !enif_get_int(env, argv[8], &want_cos))
    return enif_make_badarg(env);
  if (slot < 0 || slot >= QC_PIPELINE_SLOTS || b < 1 || n < 1 || t < 1 || tok.size != (size_t)b * n * t * 4 ||
      lens.size != (size_t)b * n * 4 || thr.size != (size_t)b * 4)
    return err(env, QC_ERR_INVALID_INPUT);
  enif_mutex_lock(r->lock);
  slot_res* sl = &r->slots[slot];
  int st = QC_ERR_INVALID_INPUT;                   /* slot still in flight */
  if (!sl->busy) {
    st = slot_prepare(sl, b, n, want_cos) ? qc_decide_batch_submit(r->ctx, slot, (const int32_t*)tok.data,
                                                                   (const int32_t*)lens.data, b, n, t,
                                                                   (const float*)thr.data, sl->ok, sl->med,
                                                                   want_cos ? sl->cos : NULL)
                                          : QC_ERR_OUT_OF_MEMORY;
    if (st == QC_OK) sl->busy = 1;
  }
  enif_mutex_unlock(r->lock);
  return st == QC_OK ? enif_make_atom(env, "ok") : err(env, st);
}

/* decide_ragged_submit(ctx, slot, tokens, offsets, n, thr, want_cos) -> :ok | {:error, atom} */
static ERL_NIF_TERM nif_decide_ragged_submit(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary tok, off, thr;
  int slot, n, want_cos;
  if (argc != 7 || !get_res(env, argv[0], &r) || !enif_get_int(env, argv[1], &slot) ||
      !enif_inspect_binary(env, argv[2], &tok) || !enif_inspect_binary(env, argv[3], &off) ||
      !enif_get_int(env, argv[4], &n) || !enif_inspect_binary(env, argv[5], &thr) || !enif_get_int(env, argv[6], &want_cos))
    return enif_make_badarg(env);
  if (slot < 0 || slot >= QC_PIPELINE_SLOTS || n < 1 || off.size < 4 || off.size % 4 != 0 || tok.size % 4 != 0)
    return err(env, QC_ERR_INVALID_INPUT);
  const int rows = (int)(off.size / 4) - 1;
  if (rows < n || rows % n != 0) return err(env, QC_ERR_INVALID_INPUT);
  const int b = rows / n;
  const int32_t* offsets = (const int32_t*)off.data;
  if (thr.size != (size_t)b * 4 || offsets[rows] < 0 || (size_t)offsets[rows] * 4 > tok.size)
    return err(env, QC_ERR_INVALID_INPUT);
  enif_mutex_lock(r->lock);
  slot_res* sl = &r->slots[slot];
  int st = QC_ERR_INVALID_INPUT;
  if (!sl->busy) {
    st = slot_prepare(sl, b, n, want_cos) ? qc_decide_ragged_submit(r->ctx, slot, (const int32_t*)tok.data, offsets, b, n,
                                                                    (const float*)thr.data, sl->ok, sl->med,
                                                                    want_cos ? sl->cos : NULL)
                                          : QC_ERR_OUT_OF_MEMORY;
    if (st == QC_OK) sl->busy = 1;
  }
  enif_mutex_unlock(r->lock);
  return st == QC_OK ? enif_make_atom(env, "ok") : err(env, st);
}

/* decide_batch_wait(ctx, slot) -> {:ok, ref_ok, medoid, cos | nil} | {:error, atom} */
static ERL_NIF_TERM nif_decide_batch_wait(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  int slot;
  if (argc != 2 || !get_res(env, argv[0], &r) || !enif_get_int(env, argv[1], &slot)) return enif_make_badarg(env);
  if (slot < 0 || slot >= QC_PIPELINE_SLOTS) return err(env, QC_ERR_INVALID_INPUT);
  enif_mutex_lock(r->lock);
  slot_res* sl = &r->slots[slot];
  if (!sl->busy) {
    enif_mutex_unlock(r->lock);
    return err(env, QC_ERR_INVALID_INPUT);
  }
  const int st = qc_decide_batch_wait(r->ctx, slot);
  sl->busy = 0;
  ERL_NIF_TERM out;
  if (st != QC_OK) {
    out = err(env, st);
  } else {
    ERL_NIF_TERM ok_t, med_t, cos_t = enif_make_atom(env, "nil");
    memcpy(enif_make_new_binary(env, (size_t)sl->b * 4, &ok_t), sl->ok, (size_t)sl->b * 4);
    memcpy(enif_make_new_binary(env, (size_t)sl->b * 4, &med_t), sl->med, (size_t)sl->b * 4);
    if (sl->want_cos) {
      const size_t cb = (size_t)sl->b * sl->n * sl->n * 4;
      memcpy(enif_make_new_binary(env, cb, &cos_t), sl->cos, cb);
    }
    out = enif_make_tuple4(env, enif_make_atom(env, "ok"), ok_t, med_t, cos_t);
  }
  enif_mutex_unlock(r->lock);
  return out;
}

/* ---- device-resident lesson rows (qc_lesson_cache_*; LessonManager dedup, lesson_manager.ex:150-173) --------------
 * A cache resource keeps its context resource alive and shares its lock. */
static ErlNifResourceType* QC_LC_TYPE = NULL;

typedef struct {
  qc_lesson_cache* lc;
  ctx_res* owner;
  int dim;
} lc_res;

static void lc_dtor(ErlNifEnv* env, void* obj) {
  (void)env;
  lc_res* r = (lc_res*)obj;
  if (r->lc) {
    enif_mutex_lock(r->owner->lock);
    qc_lesson_cache_destroy(r->lc);
    enif_mutex_unlock(r->owner->lock);
  }
  r->lc = NULL;
  if (r->owner) enif_release_resource(r->owner);
  r->owner = NULL;
}

/* lesson_cache_create(ctx, capacity, dim) -> {:ok, cache} | {:error, atom} */
static ERL_NIF_TERM nif_lesson_cache_create(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* cr;
  int capacity, dim;
  if (argc != 3 || !get_res(env, argv[0], &cr) || !enif_get_int(env, argv[1], &capacity) || !enif_get_int(env, argv[2], &dim))
    return enif_make_badarg(env);
  qc_lesson_cache* lc = NULL;
  enif_mutex_lock(cr->lock);
  int st = qc_lesson_cache_create(cr->ctx, capacity, dim, &lc);
  enif_mutex_unlock(cr->lock);
  if (st != QC_OK) return err(env, st);
  lc_res* r = (lc_res*)enif_alloc_resource(QC_LC_TYPE, sizeof(lc_res));
  r->lc = lc;
  r->owner = cr;
  r->dim = dim;
  enif_keep_resource(cr);
  ERL_NIF_TERM term = enif_make_resource(env, r);
  enif_release_resource(r);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), term);
}

static int get_lc(ErlNifEnv* env, ERL_NIF_TERM t, lc_res** out) {
  return enif_get_resource(env, t, QC_LC_TYPE, (void**)out) && (*out)->lc != NULL;
}

/* lesson_cache_put(cache, index, rows :: binary (k*dim f32)) -> :ok | {:error, atom}   (index == count appends) */
static ERL_NIF_TERM nif_lesson_cache_put(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  lc_res* r;
  ErlNifBinary rows;
  int index;
  if (argc != 3 || !get_lc(env, argv[0], &r) || !enif_get_int(env, argv[1], &index) || !enif_inspect_binary(env, argv[2], &rows))
    return enif_make_badarg(env);
  const size_t rowb = (size_t)r->dim * 4;
  if (rows.size == 0 || rows.size % rowb != 0) return err(env, QC_ERR_INVALID_INPUT);
  enif_mutex_lock(r->owner->lock);
  int st = qc_lesson_cache_put(r->lc, index, (const float*)rows.data, (int)(rows.size / rowb));
  enif_mutex_unlock(r->owner->lock);
  return st == QC_OK ? enif_make_atom(env, "ok") : err(env, st);
}

/* lesson_cache_gather(cache, src_index :: binary (n int32)) -> :ok | {:error, atom}   (prune / reorder) */
static ERL_NIF_TERM nif_lesson_cache_gather(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  lc_res* r;
  ErlNifBinary idx;
  if (argc != 2 || !get_lc(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &idx)) return enif_make_badarg(env);
  if (idx.size % 4 != 0) return err(env, QC_ERR_INVALID_INPUT);
  enif_mutex_lock(r->owner->lock);
  int st = qc_lesson_cache_gather(r->lc, (const int32_t*)idx.data, (int)(idx.size / 4));
  enif_mutex_unlock(r->owner->lock);
  return st == QC_OK ? enif_make_atom(env, "ok") : err(env, st);
}

/* lesson_cache_match(cache, query :: binary (dim f32), thr :: binary (1 f32), want_cos :: 0 | 1)
 *   -> {:ok, index :: integer (-1 = no match), cos :: binary | nil} | {:error, atom} */
static ERL_NIF_TERM nif_lesson_cache_match(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  lc_res* r;
  ErlNifBinary q, thr;
  int want_cos;
  if (argc != 4 || !get_lc(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &q) ||
      !enif_inspect_binary(env, argv[2], &thr) || !enif_get_int(env, argv[3], &want_cos))
    return enif_make_badarg(env);
  if (q.size != (size_t)r->dim * 4 || thr.size != 4) return err(env, QC_ERR_INVALID_INPUT);
  float th;
  memcpy(&th, thr.data, 4);
  enif_mutex_lock(r->owner->lock);
  const int count = qc_lesson_cache_count(r->lc);
  ERL_NIF_TERM cos_t = enif_make_atom(env, "nil");
  float* cos = (want_cos && count > 0) ? (float*)enif_make_new_binary(env, (size_t)count * 4, &cos_t) : NULL;
  int32_t index = -1;
  int st = qc_lesson_cache_match(r->lc, (const float*)q.data, th, &index, cos);
  enif_mutex_unlock(r->owner->lock);
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), enif_make_int(env, index), cos_t);
}

static ERL_NIF_TERM nif_launch_count(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  qc_ctx* ctx;
  if (argc != 1 || !get_ctx(env, argv[0], &ctx)) return enif_make_badarg(env);
  return enif_make_uint64(env, qc_launch_count(ctx));
}

static int on_load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info) {
  (void)priv;
  (void)info;
  QC_CTX_TYPE = enif_open_resource_type(env, NULL, "qc_ctx", ctx_dtor, ERL_NIF_RT_CREATE, NULL);
  QC_LC_TYPE = enif_open_resource_type(env, NULL, "qc_lesson_cache", lc_dtor, ERL_NIF_RT_CREATE, NULL);
  return (QC_CTX_TYPE == NULL || QC_LC_TYPE == NULL) ? -1 : 0;
}

static int on_upgrade(ErlNifEnv* env, void** priv, void** old_priv, ERL_NIF_TERM info) {
  (void)old_priv;
  return on_load(env, priv, info);
}

static ErlNifFunc nif_funcs[] = {
    {"ctx_create", 5, nif_ctx_create, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"decide_batch", 8, nif_decide_batch, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"embed_batch", 5, nif_embed_batch, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"embed_ragged", 3, nif_embed_ragged, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"decide_ragged", 6, nif_decide_ragged, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"similarity_batch", 6, nif_similarity_batch, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"first_match_batch", 7, nif_first_match_batch, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"chunk_mean", 5, nif_chunk_mean, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"decide_batch_submit", 9, nif_decide_batch_submit, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"decide_ragged_submit", 7, nif_decide_ragged_submit, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"decide_batch_wait", 2, nif_decide_batch_wait, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"lesson_cache_create", 3, nif_lesson_cache_create, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"lesson_cache_put", 3, nif_lesson_cache_put, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"lesson_cache_gather", 2, nif_lesson_cache_gather, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"lesson_cache_match", 4, nif_lesson_cache_match, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"launch_count", 1, nif_launch_count, 0},
};

ERL_NIF_INIT(Elixir.Quoracle.Native, nif_funcs, on_load, NULL, on_upgrade, NULL)
