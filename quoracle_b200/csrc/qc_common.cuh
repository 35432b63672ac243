// qc_common.cuh -- launch descriptors shared between the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qc {

struct EncoderLaunch {
  const CUtensorMap* w_tmap;   // host copy, passed by value as a __grid_constant__ parameter
  const uint16_t* table;       // device, bf16 [vocab x dim]
  const int32_t* token_ids;    // device [R x T]
  const int32_t* lens;         // device [R]
  float* part;                 // device [R x ceil(T/TB) x dim]
  int32_t* err_flag;           // device int32[8]: [0..3] watchdog record, [4] token-range flag
  int R, T, dim, vocab, cta_group, num_sms;
  cudaStream_t stream;
};

// K1
cudaError_t launch_encode(const EncoderLaunch& a);
int encoder_token_block(int dim);
int encoder_w_stages(int dim, int cta_group);
bool encoder_supports(int dim, int cta_group);

// K2: emb[r, :] = L2normalise( (sum_j part[r, j, :]) / len[r] ), j < ceil(len/TB); len == 0 -> zeros.
cudaError_t launch_finalize(const float* part, const int32_t* lens, float* emb, int R, int bpc, int TB, int dim,
                            cudaStream_t stream);

// K3: per decision: N x N cosine matrix, reference verdict, medoid vote.
cudaError_t launch_similarity(const float* emb, int B, int N, int D, const float* thr, int32_t* ref_ok,
                              int32_t* medoid, float* cos, int num_sms, cudaStream_t stream);

// Lesson dedup: first m with cos(query_q, rows[q, m]) >= thr[q].
cudaError_t launch_first_match(const float* query, const float* rows, int Q, int M, int D, const float* thr,
                               int32_t* out_index, float* out_cos, cudaStream_t stream);

// Unweighted chunk mean.
cudaError_t launch_chunk_mean(const float* chunks, int G, int C, int D, float* out, cudaStream_t stream);

}  // namespace qc
