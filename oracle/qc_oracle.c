/*
 * qc_oracle.c -- CPU restatement (plain C) of the arithmetic on the consensus hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under quoracle_b200/ may link, import or call this file;
 * it is the checker (tests/, __graft_entry__.smoke()) and the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference), never the product.
 *
 * Parity status: the cosine / threshold rule / chunk mean below are pinned by the reference's
 * own known-answer tests (tests/golden/reference_vectors.json, transcribed from
 * test/quoracle/actions/consensus_rules_test.exs:240-301, test/quoracle/consensus/aggregator_test.exs:878-950,
 * test/quoracle/agent/lesson_manager_test.exs:34-76).  The encoder is THIS BUILD's definition
 * (the reference calls a remote provider, lib/quoracle/models/embeddings.ex:277-298): its values
 * are "parity unpinned" against the reference and pinned only against the numpy restatement in
 * oracle/encoder.py.
 *
 * The reference itself (Elixir) cannot be compiled or run in this image (no BEAM), so there is
 * no oracle/_ref build; see DESIGN.md.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ConsensusRules.cosine_similarity/2, lib/quoracle/actions/consensus_rules.ex:193-210:
 * ONE left-to-right pass accumulating {dot, |a|^2, |b|^2} in f64; 0.0 if either norm is 0.0. */
double qo_cosine_rules(const double* a, const double* b, int d) {
  double dot = 0.0, m1 = 0.0, m2 = 0.0;
  for (int i = 0; i < d; ++i) {
    dot = dot + a[i] * b[i];
    m1 = m1 + a[i] * a[i];
    m2 = m2 + b[i] * b[i];
  }
  const double mag1 = sqrt(m1), mag2 = sqrt(m2);
  if (mag1 == 0.0 || mag2 == 0.0) return 0.0;
  return dot / (mag1 * mag2);
}

/* Aggregator.cosine_similarity/2, lib/quoracle/consensus/aggregator.ex:335-348: three passes. */
double qo_cosine_aggregator(const double* a, const double* b, int d) {
  double dot = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < d; ++i) dot = dot + a[i] * b[i];
  for (int i = 0; i < d; ++i) s1 = s1 + a[i] * a[i];
  for (int i = 0; i < d; ++i) s2 = s2 + b[i] * b[i];
  const double mag1 = sqrt(s1), mag2 = sqrt(s2);
  if (mag1 == 0.0 || mag2 == 0.0) return 0.0;
  return dot / (mag1 * mag2);
}

/* One decision: reference rule (consensus_rules.ex:277-286: all i>=1, cos(e_0,e_i) >= thr,
 * short-circuiting like Enum.all?/2) and, when want_full != 0, the full N x N matrix + medoid
 * (north-star vote, SURVEY.md 8(a) A13: argmax_i mean_j cos(e_i,e_j), first maximum wins).
 * emb is f64 [N x D]; cos_out (optional) [N x N]. */
void qo_decide_one(const double* emb, int N, int D, double thr, int want_full, int* ref_ok, int* medoid,
                   double* cos_out) {
  int ok = 1;
  for (int i = 1; i < N; ++i) {
    if (!(qo_cosine_rules(emb, emb + (size_t)i * D, D) >= thr)) {
      ok = 0;
      break;
    }
  }
  if (ref_ok) *ref_ok = ok;
  if (!want_full) return;
  double* c = cos_out ? cos_out : (double*)malloc(sizeof(double) * N * N);
  for (int i = 0; i < N; ++i)
    for (int j = i; j < N; ++j) {
      const double v = qo_cosine_rules(emb + (size_t)i * D, emb + (size_t)j * D, D);
      c[i * N + j] = v;
      c[j * N + i] = v;
    }
  int best = 0;
  double best_mean = 0.0;
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int j = 0; j < N; ++j) s = s + c[i * N + j];
    const double m = s / (double)N;
    if (i == 0 || m > best_mean) {
      best_mean = m;
      best = i;
    }
  }
  if (medoid) *medoid = best;
  if (!cos_out) free(c);
}

/* Batch over B decisions; emb f32 [B x N x D] widened to f64 (what BEAM would hold after the NIF
 * hands back floats); thr f64 [B].  threads <= 1: strictly serial (faithful to the reference, which
 * is serial within and across the values of a decision); otherwise one decision per OpenMP thread. */
void qo_similarity_batch(const float* emb, int B, int N, int D, const double* thr, int want_full, int threads,
                         int32_t* ref_ok, int32_t* medoid, double* cos_out) {
#ifdef _OPENMP
  if (threads > 1) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) if (threads > 1)
  for (int b = 0; b < B; ++b) {
    double* e = (double*)malloc(sizeof(double) * (size_t)N * D);
    const float* src = emb + (size_t)b * N * D;
    for (size_t i = 0; i < (size_t)N * D; ++i) e[i] = (double)src[i];
    int ok = 0, med = 0;
    qo_decide_one(e, N, D, thr[b], want_full, &ok, &med, cos_out ? cos_out + (size_t)b * N * N : NULL);
    if (ref_ok) ref_ok[b] = ok;
    if (medoid && want_full) medoid[b] = med;
    free(e);
  }
}

/* Embeddings.average_embeddings/1, lib/quoracle/models/embeddings.ex:371-383: per element, sum the
 * chunks in order starting from integer 0, divide by the chunk count; NOT re-normalised. */
void qo_chunk_mean(const double* chunks, int C, int D, double* out) {
  for (int d = 0; d < D; ++d) {
    double s = 0.0;
    for (int c = 0; c < C; ++c) s = s + chunks[(size_t)c * D + d];
    out[d] = s / (double)C;
  }
}

/* LessonManager.find_similar_lesson/5, lib/quoracle/agent/lesson_manager.ex:150-173 (arithmetic part):
 * first m with Aggregator.cosine_similarity(query, rows[m]) >= thr, else -1. */
int qo_first_match(const double* query, const double* rows, int M, int D, double thr) {
  for (int m = 0; m < M; ++m)
    if (qo_cosine_aggregator(query, rows + (size_t)m * D, D) >= thr) return m;
  return -1;
}

/* ---- encoder (this build's definition) -------------------------------------------- */

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, sizeof(f));
  return f;
}

static inline double gelu_tanh_f64(double x) {
  return 0.5 * x * (1.0 + tanh(0.7978845608028654 * (x + 0.044715 * x * x * x)));
}

/* Exact-ish version: f64 accumulation, libm tanh.  h = L2norm(mean_t GELU(E[tok_t] . W)); len 0 -> zeros.
 * table bf16 [vocab x D], w bf16 [D x D] (in x out), token_ids [R x T], lens [R], out f64 [R x D]. */
void qo_encode_f64(const uint16_t* table, const uint16_t* w, int vocab, int D, const int32_t* token_ids,
                   const int32_t* lens, int R, int T, int threads, double* out) {
  double* wd = (double*)malloc(sizeof(double) * (size_t)D * D);
  for (size_t i = 0; i < (size_t)D * D; ++i) wd[i] = (double)bf16_to_f32(w[i]);
#ifdef _OPENMP
  if (threads > 1) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) if (threads > 1)
  for (int r = 0; r < R; ++r) {
    double* pooled = out + (size_t)r * D;
    double* y = (double*)malloc(sizeof(double) * D);
    double* x = (double*)malloc(sizeof(double) * D);
    for (int d = 0; d < D; ++d) pooled[d] = 0.0;
    const int len = lens[r];
    for (int t = 0; t < len; ++t) {
      const int32_t id = token_ids[(size_t)r * T + t];
      if (id < 0 || id >= vocab) continue; /* pad marker inside len: contributes GELU(0) = 0 */
      const uint16_t* e = table + (size_t)id * D;
      for (int k = 0; k < D; ++k) x[k] = (double)bf16_to_f32(e[k]);
      for (int d = 0; d < D; ++d) y[d] = 0.0;
      for (int k = 0; k < D; ++k) {
        const double a = x[k];
        const double* wr = wd + (size_t)k * D;
        for (int d = 0; d < D; ++d) y[d] += a * wr[d];
      }
      for (int d = 0; d < D; ++d) pooled[d] += gelu_tanh_f64(y[d]);
    }
    double ss = 0.0;
    for (int d = 0; d < D; ++d) {
      pooled[d] = len > 0 ? pooled[d] / (double)len : 0.0;
      ss += pooled[d] * pooled[d];
    }
    const double nrm = sqrt(ss);
    for (int d = 0; d < D; ++d) pooled[d] = nrm > 0.0 ? pooled[d] / nrm : 0.0;
    free(y);
    free(x);
  }
  free(wd);
}

/* Timed CPU baseline version: f32, 4-token register/L1 blocking, vectorisable inner loops, tanhf.
 * Same function as qo_encode_f64 up to f32 rounding.  out f32 [R x D]. */
__attribute__((optimize("fp-contract=fast")))
void qo_encode_f32(const uint16_t* table, const uint16_t* w, int vocab, int D, const int32_t* token_ids,
                   const int32_t* lens, int R, int T, int threads, float* out) {
  float* wf = (float*)aligned_alloc(64, sizeof(float) * (size_t)D * D);
  for (size_t i = 0; i < (size_t)D * D; ++i) wf[i] = bf16_to_f32(w[i]);
#ifdef _OPENMP
  if (threads > 1) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) if (threads > 1)
  for (int r = 0; r < R; ++r) {
    enum { TBLK = 4 };
    float* acc = (float*)aligned_alloc(64, sizeof(float) * TBLK * (size_t)D);
    float* x = (float*)aligned_alloc(64, sizeof(float) * TBLK * (size_t)D);
    float* pooled = (float*)aligned_alloc(64, sizeof(float) * (size_t)D);
    for (int d = 0; d < D; ++d) pooled[d] = 0.f;
    const int len = lens[r];
    for (int t0 = 0; t0 < len; t0 += TBLK) {
      const int nt = len - t0 < TBLK ? len - t0 : TBLK;
      for (int q = 0; q < TBLK; ++q) {
        int32_t id = q < nt ? token_ids[(size_t)r * T + t0 + q] : -1;
        if (id >= vocab) id = -1;
        for (int k = 0; k < D; ++k) x[(size_t)q * D + k] = id >= 0 ? bf16_to_f32(table[(size_t)id * D + k]) : 0.f;
      }
      memset(acc, 0, sizeof(float) * TBLK * (size_t)D);
      for (int k = 0; k < D; ++k) {
        const float* wr = wf + (size_t)k * D;
        const float a0 = x[k], a1 = x[(size_t)D + k], a2 = x[2 * (size_t)D + k], a3 = x[3 * (size_t)D + k];
        float* c0 = acc;
        float* c1 = acc + D;
        float* c2 = acc + 2 * (size_t)D;
        float* c3 = acc + 3 * (size_t)D;
#pragma omp simd
        for (int d = 0; d < D; ++d) {
          const float wv = wr[d];
          c0[d] += a0 * wv;
          c1[d] += a1 * wv;
          c2[d] += a2 * wv;
          c3[d] += a3 * wv;
        }
      }
      for (int q = 0; q < nt; ++q) {
        const float* c = acc + (size_t)q * D;
        for (int d = 0; d < D; ++d) {
          const float v = c[d];
          pooled[d] += 0.5f * v * (1.f + tanhf(0.7978845608f * (v + 0.044715f * v * v * v)));
        }
      }
    }
    float ss = 0.f;
    for (int d = 0; d < D; ++d) {
      pooled[d] = len > 0 ? pooled[d] / (float)len : 0.f;
      ss += pooled[d] * pooled[d];
    }
    const float nrm = sqrtf(ss);
    for (int d = 0; d < D; ++d) out[(size_t)r * D + d] = nrm > 0.f ? pooled[d] / nrm : 0.f;
    free(acc);
    free(x);
    free(pooled);
  }
  free(wf);
}

int qo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
