/*
 * qconsensus.h -- C ABI of the B200-native consensus hot path (libqconsensus.so).
 *
 * This is the drop-in boundary for the one compute path of shelvick/quoracle that
 * this repository accelerates: embed the N candidate values of one action
 * parameter, cosine-compare them, and emit the verdict.  The reference has no FFI
 * on this path (it is pure Elixir); the seams an Erlang NIF built on this header
 * plugs into are function-valued options.  Each entry point cites the reference
 * interface it stands behind (paths relative to the reference checkout):
 *
 *   qc_embed_batch        <- `:embedding_fn` 2-arity closure,
 *                            lib/quoracle/actions/consensus_rules.ex:302-327, and
 *                            Quoracle.Models.Embeddings.get_embedding/2,
 *                            lib/quoracle/models/embeddings.ex:41-63 (the provider
 *                            HTTP call at :277-298 is what the local encoder replaces)
 *   qc_similarity_batch   <- ConsensusRules.cosine_similarity/2 + the
 *                            `all i>=1: cos(e_0,e_i) >= threshold` rule,
 *                            lib/quoracle/actions/consensus_rules.ex:189-210,277-286;
 *                            Aggregator.cosine_similarity/2,
 *                            lib/quoracle/consensus/aggregator.ex:335-350
 *   qc_decide_batch       <- use_embeddings_similarity/5 for every pending decision
 *                            of the agent tree at once,
 *                            lib/quoracle/actions/consensus_rules.ex:249-299
 *                            (per-agent caller: lib/quoracle/agent/consensus_handler.ex:64)
 *   qc_first_match_batch  <- LessonManager.find_similar_lesson/5,
 *                            lib/quoracle/agent/lesson_manager.ex:150-173
 *   qc_chunk_mean         <- Embeddings.average_embeddings/1,
 *                            lib/quoracle/models/embeddings.ex:371-383
 *
 * Conventions
 *   - Plain C: pointers + sizes, no C++/torch types.  Host entry points take HOST
 *     memory owned by the caller; `_dev` entry points take DEVICE memory on the
 *     context's device and a cudaStream_t passed as void* (NULL = context stream).
 *   - Return value: QC_OK (0) or a negative qc_status.  Nothing is thrown.  Each
 *     status maps 1:1 to the Elixir error atom the reference uses on this path
 *     (lib/quoracle/models/embeddings.ex:49-50,184,228-229,300-331,344).
 *   - A context is bound to ONE device (one process per GPU); calls on one context
 *     are serialised by the caller (one host thread at a time).  The WORK of `_dev`
 *     calls enqueued on different streams may overlap on the GPU: each caller stream
 *     (up to four at a time) gets its own K1->K2 partial-sum and embedding scratch.
 *     No globals, no CPU fallback: every entry point
 *     fails with QC_ERR_NO_DEVICE when no sm_100 device is usable.
 *   - Zero-norm vectors give cosine 0.0 (never NaN), as both reference cosines do
 *     (consensus_rules.ex:205-209, aggregator.ex:343-347).
 */
#ifndef QCONSENSUS_H
#define QCONSENSUS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QC_ABI_VERSION 1

typedef enum qc_status {
  QC_OK = 0,
  QC_ERR_INVALID_INPUT = -1,      /* :invalid_input          */
  QC_ERR_NOT_CONFIGURED = -2,     /* :not_configured         (no encoder weights loaded) */
  QC_ERR_NO_DEVICE = -3,          /* :service_unavailable    (no usable sm_100 GPU)      */
  QC_ERR_CUDA = -4,               /* :service_unavailable    (CUDA runtime error)        */
  QC_ERR_OUT_OF_MEMORY = -5,      /* :service_unavailable                                 */
  QC_ERR_CONTEXT_LENGTH = -6,     /* :context_length_exceeded (T > encoder max_tokens)    */
  QC_ERR_UNSUPPORTED = -7,        /* :bad_request            (shape outside the ABI)      */
  QC_ERR_TOKEN_RANGE = -8,        /* :invalid_input          (token id >= vocab)          */
  QC_ERR_KERNEL_TIMEOUT = -9      /* :service_unavailable    (device barrier watchdog)    */
} qc_status;

/* Limits of this ABI version. */
#define QC_MAX_POOL 64            /* N: candidates per decision (BASELINE sweep: 2..64)   */
#define QC_PAD_TOKEN (-1)         /* padding marker in token_ids                          */
#define QC_MAX_TOKENS 8192        /* T: token slots per candidate; the reference chunks above
                                     floor(0.9*8191) tokens, lib/quoracle/models/embeddings/token_chunker.ex:13-15,48-53 */

/*
 * The local encoder (the reference fixes none -- SURVEY.md section 8(d)):
 *   h = L2normalise( mean_{t < len} GELU_tanh( E[tok_t] . W ) )
 * E: bf16 [vocab x dim] row-major, W: bf16 [dim x dim] row-major (in x out),
 * fp32 accumulation.  dim must be one of 384 / 768 / 1536.
 * Both arrays are HOST pointers to raw bf16 bit patterns; copied at create time.
 */
typedef struct qc_encoder_cfg {
  int32_t vocab;
  int32_t dim;
  const uint16_t* table_bf16; /* [vocab * dim] */
  const uint16_t* w_bf16;     /* [dim * dim]   */
} qc_encoder_cfg;

typedef struct qc_ctx qc_ctx;

/* Context --------------------------------------------------------------------------- */

/* enc may be NULL: similarity-only context (embed/decide then return QC_ERR_NOT_CONFIGURED). */
int qc_ctx_create(qc_ctx** out, int device_id, const qc_encoder_cfg* enc);
void qc_ctx_destroy(qc_ctx* ctx);
int qc_ctx_device(const qc_ctx* ctx);
int qc_ctx_dim(const qc_ctx* ctx);
/* Blocks until all work queued on the context stream has finished. */
int qc_ctx_synchronize(qc_ctx* ctx);

/* Pinned host memory (so a NIF can stage binaries without an extra copy). */
void* qc_host_alloc(size_t bytes);
void qc_host_free(void* p);

/* Host entry points ----------------------------------------------------------------- */

/*
 * Full path for B decisions of N candidates each, T token slots per candidate.
 *   token_ids [B*N*T] int32, QC_PAD_TOKEN beyond lens; lens [B*N] int32 in [0,T];
 *   thr [B] f32 (the schema threshold of the parameter being merged).
 * Outputs (each may be NULL):
 *   out_ref_ok [B]  1 iff for all i>=1: cos(e_0,e_i) >= thr  (reference rule: value_0 wins)
 *   out_medoid [B]  argmax_i mean_j cos(e_i,e_j), lowest index on ties (north-star vote)
 *   out_cos    [B*N*N] f32 cosine matrix,  out_emb [B*N*dim] f32 normalised embeddings.
 */
int qc_decide_batch(qc_ctx* ctx, const int32_t* token_ids, const int32_t* lens, int B, int N, int T,
                    const float* thr, int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos,
                    float* out_emb);

/*
 * Pipelined form of qc_decide_batch: `submit` enqueues the copies and kernels of one batch on one of
 * QC_PIPELINE_SLOTS slots (each with its own stream and workspaces) and returns; `wait` blocks until that slot's
 * outputs are in the caller's buffers.  A caller (the NIF collector) that keeps two batches in flight overlaps the
 * host->device copy of one with the GPU work of the other; a third keeps an encoder launch queued while the
 * host turns a finished slot around (the small K2/K3 kernels of a finished batch cannot always co-reside with
 * the next batch's encoder, which holds 225 KB of shared memory per SM).  Input and output buffers must stay
 * valid and untouched until `wait` returns.
 */
#define QC_PIPELINE_SLOTS 4
int qc_decide_batch_submit(qc_ctx* ctx, int slot, const int32_t* token_ids, const int32_t* lens, int B, int N,
                           int T, const float* thr, int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos);
int qc_decide_batch_wait(qc_ctx* ctx, int slot);

/*
 * Ragged (CSR) forms of qc_embed_batch / qc_decide_batch for callers whose candidates differ in length (what a
 * collector of real model responses holds): tokens of candidate r are tokens[offsets[r] .. offsets[r+1]), offsets has
 * R + 1 (= B*N + 1) non-decreasing entries starting anywhere >= 0, no candidate longer than QC_MAX_TOKENS.  Nothing is
 * padded on the host or on the wire; on the device the candidates are cut into chunks and laid end to end (packed
 * token stream), so the encoder's work is proportional to the tokens that exist.  Same results as the padded entry
 * points up to fp32 summation order of the pooled sum.
 */
int qc_embed_ragged(qc_ctx* ctx, const int32_t* tokens, const int32_t* offsets, int R, float* out_emb);
int qc_decide_ragged(qc_ctx* ctx, const int32_t* tokens, const int32_t* offsets, int B, int N, const float* thr,
                     int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos, float* out_emb);
/* Pipelined form (slots and qc_decide_batch_wait as for qc_decide_batch_submit; buffers stay untouched until wait). */
int qc_decide_ragged_submit(qc_ctx* ctx, int slot, const int32_t* tokens, const int32_t* offsets, int B, int N,
                            const float* thr, int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos);

/* Embedding only: R candidate rows -> out_emb [R*dim] f32 (what an `:embedding_fn` returns). */
int qc_embed_batch(qc_ctx* ctx, const int32_t* token_ids, const int32_t* lens, int R, int T,
                   float* out_emb);

/* Similarity only, for callers that already hold embeddings: emb [B*N*D] f32 (any D >= 1). */
int qc_similarity_batch(qc_ctx* ctx, const float* emb, int B, int N, int D, const float* thr,
                        int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos);

/*
 * Lesson dedup: for each of Q queries, the FIRST row m of its M candidates with
 * cos(query, row_m) >= thr[q], else -1.  query [Q*D], rows [Q*M*D] f32.
 * out_cos (optional) [Q*M].
 */
int qc_first_match_batch(qc_ctx* ctx, const float* query, const float* rows, int Q, int M, int D,
                         const float* thr, int32_t* out_index, float* out_cos);

/*
 * Device-resident lesson rows for the dedup above.  The reference re-embeds every existing
 * lesson (<= 100 per agent, lesson_manager.ex:150-173) for every new lesson; with a cache the rows of one agent's
 * lesson list stay in HBM in LIST ORDER and a dedup call uploads one query row:
 *   put     rows [nrows*dim] f32 (host) are written at index .. index+nrows-1; index <= count (index == count appends;
 *           a merged lesson REPLACES its row: the merged lesson keeps the NEW content, lesson_manager.ex:159-161)
 *   gather  the list was pruned / reordered: new row i = old row src_index[i], count becomes n
 *   match   first resident row m (ascending) with cos(query, row_m) >= thr, else -1; out_cos (optional) [count]
 * Calls are serialised with the other calls of the owning context.  capacity is fixed at creation.
 */
typedef struct qc_lesson_cache qc_lesson_cache;
int qc_lesson_cache_create(qc_ctx* ctx, int capacity, int dim, qc_lesson_cache** out);
void qc_lesson_cache_destroy(qc_lesson_cache* cache);
int qc_lesson_cache_count(const qc_lesson_cache* cache);
int qc_lesson_cache_put(qc_lesson_cache* cache, int index, const float* rows, int nrows);
int qc_lesson_cache_gather(qc_lesson_cache* cache, const int32_t* src_index, int n);
int qc_lesson_cache_match(qc_lesson_cache* cache, const float* query, float thr, int32_t* out_index, float* out_cos);

/* Unweighted, un-normalised element-wise mean of C chunk vectors per group: in [G*C*D] -> out [G*D]. */
int qc_chunk_mean(qc_ctx* ctx, const float* chunks, int G, int C, int D, float* out_mean);

/* Device entry points (inputs/outputs already resident in HBM) ------------------------ */

int qc_decide_batch_dev(qc_ctx* ctx, const int32_t* d_token_ids, const int32_t* d_lens, int B, int N,
                        int T, const float* d_thr, int32_t* d_ref_ok, int32_t* d_medoid,
                        float* d_cos, float* d_emb, void* stream);
int qc_embed_batch_dev(qc_ctx* ctx, const int32_t* d_token_ids, const int32_t* d_lens, int R, int T,
                       float* d_emb, void* stream);
int qc_similarity_batch_dev(qc_ctx* ctx, const float* d_emb, int B, int N, int D, const float* d_thr,
                            int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos, void* stream);

/*
 * CUDA graphs over the device entry points.  Between qc_capture_begin and qc_capture_end every `_dev` call made on
 * `stream` (a non-default stream) is RECORDED instead of executed; qc_capture_end returns a graph that
 * qc_graph_launch replays with one launch -- the whole inner loop of a caller whose batches are resident in HBM,
 * with no host work (and no host jitter) between the kernels.  The same calls must have run once before the capture
 * (workspaces are sized on first use and cannot grow while recording).  With qc_profile_enable on, the per-kernel
 * event brackets are recorded INTO the graph as event-record nodes: replay the graph once, then qc_kernel_stats returns
 * device-side durations with no host time inside any bracket.  Pointers passed to the recorded calls are baked into
 * the graph and must stay valid while it is replayed.
 */
typedef struct qc_graph qc_graph;
int qc_capture_begin(qc_ctx* ctx, void* stream);
int qc_capture_end(qc_ctx* ctx, void* stream, qc_graph** out);
int qc_graph_launch(qc_graph* graph, void* stream);
void qc_graph_destroy(qc_graph* graph);

/* Introspection ---------------------------------------------------------------------- */

const char* qc_strerror(int status);
/* Elixir atom name the status maps to (without the leading colon). */
const char* qc_status_atom(int status);
int qc_abi_version(void);

/* Kernel ids for the counters below. */
#define QC_K_ENCODE 0     /* K1: token gather + tcgen05 GEMM + GELU + token-sum          */
#define QC_K_FINALIZE 1   /* K2: block-sum -> mean -> L2 normalise                        */
#define QC_K_SIMILARITY 2 /* K3: N x N cosine + reference verdict + medoid vote           */
#define QC_K_FIRST_MATCH 3
#define QC_K_CHUNK_MEAN 4
#define QC_K_COUNT 5

/*
 * Synchronises the context and reports device-side status raised by earlier (asynchronous)
 * `_dev` calls: QC_ERR_TOKEN_RANGE, QC_ERR_KERNEL_TIMEOUT or QC_OK.  Host entry points call it
 * themselves.  qc_last_hang returns {flag, barrier tag, block, thread} of the last watchdog hit.
 */
int qc_check_status(qc_ctx* ctx);
int qc_last_hang(const qc_ctx* ctx, int32_t info[4]);

/* Number of kernels of this library launched on the context since the last reset. */
uint64_t qc_launch_count(const qc_ctx* ctx);
/*
 * Per-kernel device timing with CUDA events on the launching stream.  When enabled,
 * every launch is bracketed by events; qc_kernel_stats synchronises the context and
 * returns accumulated milliseconds and launch counts per kernel id since the reset.
 */
int qc_profile_enable(qc_ctx* ctx, int on);
int qc_profile_reset(qc_ctx* ctx);
int qc_kernel_stats(qc_ctx* ctx, double ms[QC_K_COUNT], uint64_t launches[QC_K_COUNT]);
/* Encoder tiling in use: cta_group (1|2), token block, W stages. Returns QC_OK / NOT_CONFIGURED. */
int qc_encoder_info(const qc_ctx* ctx, int32_t* cta_group, int32_t* token_block, int32_t* w_stages);
/*
 * Tuning aid for K1: when enabled, every encoder launch records, per CTA, the cycles each warp role
 * spent in total and blocked on its barriers; a call with out != NULL returns the per-CTA averages of
 * the LAST launch: [0] MMA issuer total, [1] waiting for W tiles, [2] waiting for gathered tokens,
 * [3] waiting for a free accumulator, [4] epilogue total, [5] waiting for accumulators, [6] gather
 * total, [7] waiting for free token buffers, [8] W producer total, [9] waiting for free W stages.
 */
int qc_encoder_stall_profile(qc_ctx* ctx, int enable, double out[16]);
/*
 * Pool split across GPUs (north star: "a single all-gather of embedding rows over NVLink only
 * when one agent's model pool spans devices").  The reference has no multi-device notion -- its pool is the
 * model_ids list of Quoracle.Consensus.get_consensus (lib/quoracle/consensus.ex) -- so this is the seam a
 * multi-GPU deployment adds: `world` processes (one per GPU of one node), each holding `pool_local` consecutive
 * members of every decision's pool (pool order = rank-major).  Every rank creates a split object, exports its
 * handle, and attaches every peer's handle (exchanged by the host runtime: Erlang distribution, or
 * torch.distributed in this repo's tests).  qc_decide_split_dev then runs K1 on the local candidates, K2 stores
 * each normalised row straight into every rank's gathered buffer through peer memory (no separate collective),
 * a device-side flag handshake orders the ranks, and K3 decides on the full pool: every rank gets all verdicts.
 * All ranks must make the same sequence of calls with the same B.  d_* pointers are device memory of the
 * context's GPU; `stream` as in the other _dev entry points.  A peer that never arrives ends the wait after ~3 s
 * and surfaces as QC_ERR_KERNEL_TIMEOUT from qc_check_status.
 */
typedef struct qc_split qc_split;
#define QC_SPLIT_HANDLE_BYTES 64
#define QC_MAX_SPLIT_WORLD 8
int qc_split_create(qc_ctx* ctx, int world, int rank, int max_decisions, int pool_local, qc_split** out);
int qc_split_export(qc_split* split, uint8_t handle[QC_SPLIT_HANDLE_BYTES]);
int qc_split_attach(qc_split* split, int peer_rank, const uint8_t handle[QC_SPLIT_HANDLE_BYTES]);
int qc_decide_split_dev(qc_ctx* ctx, qc_split* split, const int32_t* d_token_ids, const int32_t* d_lens, int B, int T,
                        const float* d_thr, int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos, void* stream);
/* Device pointer to the gathered rows [decisions, world * pool_local, dim] of the last qc_decide_split_dev. */
int qc_split_gathered(const qc_split* split, const float** d_emb, int32_t* decisions);
int qc_split_destroy(qc_split* split);

/*
 * Long candidates sharded on T across GPUs (BASELINE config 4).  Rank p holds the p-th token
 * slice of EVERY candidate: d_token_ids [R x T_local], d_lens[r] = valid tokens of row r inside this slice (0 when
 * the candidate ends earlier).  K1 runs on the slice; one fp32 partial sum and the token count per row go into
 * every rank's exchange buffer through peer memory; after the flag handshake each rank forms
 * L2normalise(sum_p partial_p / sum_p len_p) -- the unsplit encoder's value up to fp32 summation order, identical
 * on all ranks.  (The reference's own long-text path is the unweighted chunk mean of qc_chunk_mean,
 * lib/quoracle/models/embeddings.ex:173-192,371-383; this is the encoder itself spread over devices.)
 * Handles are exported / attached with qc_split_export / qc_split_attach and released with qc_split_destroy.
 * qc_decide_tsplit_dev adds K3; d_emb may be NULL (context workspace) or receives the [B*N, dim] rows.
 */
int qc_tsplit_create(qc_ctx* ctx, int world, int rank, int max_rows, qc_split** out);
int qc_embed_tsplit_dev(qc_ctx* ctx, qc_split* split, const int32_t* d_token_ids, const int32_t* d_lens, int R,
                        int T_local, float* d_emb, void* stream);
int qc_decide_tsplit_dev(qc_ctx* ctx, qc_split* split, const int32_t* d_token_ids, const int32_t* d_lens, int B, int N,
                         int T_local, const float* d_thr, int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos,
                         float* d_emb, void* stream);

/*
 * Ragged batches on the device entry points.  K1 works on token blocks of qc_encoder_token_block(T) tokens; block
 * (r, j) of candidate r has id r * ceil(T / token_block) + j and is non-empty when j * token_block < len[r].  The host
 * entry points (which hold the lengths) list the non-empty blocks themselves; a caller of the *_dev entry points can
 * do the same: upload the ascending ids of the non-empty blocks and pass them here -- the hint applies to the NEXT
 * encoder launch of this context only (qc_embed_batch_dev / qc_decide_batch_dev / the split entry points) and the
 * buffer must stay valid until that launch has run.  Without a hint every block is visited and empty ones are
 * skipped (a CTA pair whose second block is empty then runs half empty).  Results are identical either way.
 */
int qc_encoder_token_block(qc_ctx* ctx, int T, int32_t* token_block);
/*
 * Ragged batches on the host entry points (qc_decide_batch, qc_decide_batch_submit, qc_embed_batch): when it saves
 * at least 5 % of the token slots the valid tokens are packed into a stream of chunks (32 tokens, or 16 / 8 when 32
 * does not divide the token block; each chunk owned by one candidate), K1 runs on the packed stream and K2 adds a
 * candidate's chunk sums.
 * Same arithmetic as the unpacked walk up to fp32 summation order (|delta| ~ 1e-8 on unit rows).
 * mode 0 = auto (default), 1 = never pack, 2 = always pack.
 */
int qc_encoder_set_packing(qc_ctx* ctx, int mode);
/* Host-only test hook (needs no device): the packed-stream plan for `lens` at the given token block, written to
 * out as [chunk_src | chunk_n | chunk_row] (nchunks entries each, sources relative to a padded [R x T] id matrix)
 * followed by row_start (R + 1 entries); out_capacity counts int32 slots. */
int qc_debug_pack_plan(const int32_t* lens, int R, int T, int token_block, int32_t* out, int64_t out_capacity,
                       int32_t* nchunks, int32_t* nruns);
int qc_encoder_hint_block_map(qc_ctx* ctx, const int32_t* d_block_map, int nslots);

/* Test/tuning hook: K3 kernel for pools of 9..64 candidates. 0 = auto (tcgen05 kind::tf32 kernel when the
 * matrix is TMA-addressable: D % 4 == 0, 16-byte aligned base; the mma.sync kernel otherwise), 1 = always the
 * mma.sync kernel, 2 = tcgen05 only (QC_ERR_UNSUPPORTED when the layout does not allow it). */
int qc_similarity_set_path(qc_ctx* ctx, int path);
/* Test hook: force the encoder variant (cta_group 1 or 2) for subsequent launches. */
int qc_encoder_set_cta_group(qc_ctx* ctx, int cta_group);
/* Test/tuning hook: select the K1 tiling (tokens per resident block, CTAs per MMA group, W stage k-width,
 * gather mode 0 = cp.async / 1 = TMA gather4).  token_block 0 returns to the default: the tiling is
 * chosen per call from T (fastest measured tiling weighted by the fraction of non-padding token slots). */
int qc_encoder_set_variant(qc_ctx* ctx, int token_block, int cta_group, int w_bk, int gather_mode);
/* W^T multicast degree of K1: `pairs` CTA pairs of one thread-block cluster (2 -> cluster of 4 SMs, 4 -> cluster of 8)
 * walk the weight tiles in lock step, each tile read from L2 once and delivered to all of them by TMA multicast.
 * 0 = the tiling's own default, 1 = off.  Applied to every later launch whose tiling has such an instantiation
 * (CTA-pair tilings with TMA gather); other tilings ignore it.  qc_encoder_multicast reports the degree in effect. */
int qc_encoder_set_multicast(qc_ctx* ctx, int pairs);
int qc_encoder_multicast(const qc_ctx* ctx, int32_t* pairs);

#ifdef __cplusplus
}
#endif
#endif /* QCONSENSUS_H */
