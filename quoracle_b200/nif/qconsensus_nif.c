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
 * NOT BUILT IN THIS REPOSITORY'S CI: there is no BEAM / erl_nif.h in the image.  The file is type-checked
 * against nif/stub/erl_nif.h; build it where OTP is installed ("Building the NIF").
 */
#include <erl_nif.h>
#include <string.h>

#include "qconsensus.h"

static ErlNifResourceType* QC_CTX_TYPE = NULL;

typedef struct {
  qc_ctx* ctx;
} ctx_res;

static void ctx_dtor(ErlNifEnv* env, void* obj) {
  (void)env;
  ctx_res* r = (ctx_res*)obj;
  if (r->ctx) qc_ctx_destroy(r->ctx);
  r->ctx = NULL;
}

static ERL_NIF_TERM err(ErlNifEnv* env, int status) {
  return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_atom(env, qc_status_atom(status)));
}

static int get_ctx(ErlNifEnv* env, ERL_NIF_TERM t, qc_ctx** out) {
  ctx_res* r;
  if (!enif_get_resource(env, t, QC_CTX_TYPE, (void**)&r) || r->ctx == NULL) return 0;
  *out = r->ctx;
  return 1;
}

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
  r->ctx = ctx;
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
  int st = qc_decide_batch(ctx, (const int32_t*)tok.data, (const int32_t*)lens.data, b, n, t, (const float*)thr.data,
                           ref_ok, medoid, cos, NULL);
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
  int st = qc_embed_batch(ctx, (const int32_t*)tok.data, (const int32_t*)lens.data, r, t, emb);
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
  int st = qc_embed_ragged(ctx, (const int32_t*)tok.data, offsets, r, emb);
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
  int st = qc_decide_ragged(ctx, (const int32_t*)tok.data, offsets, b, n, (const float*)thr.data, ref_ok, medoid, cos, NULL);
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
  int st = qc_similarity_batch(ctx, (const float*)emb.data, b, n, d, (const float*)thr.data, ref_ok, medoid, cos);
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
  int st = qc_first_match_batch(ctx, (const float*)query.data, (const float*)rows.data, q, m, d,
                                (const float*)thr.data, idx, NULL);
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
  int st = qc_chunk_mean(ctx, (const float*)chunks.data, g, c, d, out);
  if (st != QC_OK) return err(env, st);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), out_t);
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
  return QC_CTX_TYPE == NULL ? -1 : 0;
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
    {"launch_count", 1, nif_launch_count, 0},
};

ERL_NIF_INIT(Elixir.Quoracle.Native, nif_funcs, on_load, NULL, on_upgrade, NULL)
