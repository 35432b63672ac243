// qc_common.cuh -- launch descriptors shared between the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qc {

// K1 tiling variant: tokens per resident block, CTAs per MMA group (1 | 2), k-width of a streamed W stage.
struct EncoderVariant {
  int token_block;
  int cta_group;
  int w_bk;
  int gather;   // 0 = cp.async (LDGSTS) gather, 1 = TMA gather4
};

struct EncoderLaunch {
  const CUtensorMap* w_tmap64; // W^T, box 64 x 128, 128-byte swizzle (host copy; passed __grid_constant__)
  const CUtensorMap* w_tmap32; // W^T, box 32 x 128, 64-byte swizzle
  const CUtensorMap* e_tmap;   // E [vocab x dim], box 64 x 1, 128-byte swizzle (gather4)
  const uint16_t* table;       // device, bf16 [vocab x dim]
  const int32_t* token_ids;    // device [R x T]
  const int32_t* lens;         // device [R]
  float* part;                 // device [R x ceil(T/TB) x dim]
  int32_t* err_flag;           // device int32[8]: [0..3] watchdog record, [4] token-range flag
  unsigned long long* stall;   // device [grid x 16] per-role stall cycles, or nullptr
  int R, T, dim, vocab, num_sms;
  EncoderVariant variant;
  cudaStream_t stream;
};

// K1
cudaError_t launch_encode(const EncoderLaunch& a);
bool encoder_variant_supported(int dim, const EncoderVariant& v);
int encoder_variant_w_stages(int dim, const EncoderVariant& v);
EncoderVariant encoder_default_variant(int dim);
EncoderVariant encoder_auto_variant(int dim, int T);

// K2: emb[r, :] = L2normalise( (sum_j part[r, j, :]) / len[r] ), j < ceil(len/TB); len == 0 -> zeros.
cudaError_t launch_finalize(const float* part, const int32_t* lens, float* emb, int R, int bpc, int TB, int dim,
                            cudaStream_t stream);

// K3: per decision: N x N cosine matrix, reference verdict, medoid vote.
cudaError_t launch_similarity(const float* emb, int B, int N, int D, const float* thr, int32_t* ref_ok,
                              int32_t* medoid, float* cos, int num_sms, cudaStream_t stream);

// K3 for pools of 9..64 on tcgen05 (kind::tf32, 3xTF32 split); x_tmap views emb as [B*N x D] fp32, box 32 x 128.
cudaError_t launch_similarity_tc(const CUtensorMap* x_tmap, int B, int N, int D, const float* thr, int32_t* ref_ok,
                                 int32_t* medoid, float* cos, int32_t* dbg, unsigned long long* stall, int num_sms,
                                 cudaStream_t stream);

// Lesson dedup: first m with cos(query_q, rows[q, m]) >= thr[q].
cudaError_t launch_first_match(const float* query, const float* rows, int Q, int M, int D, const float* thr,
                               int32_t* out_index, float* out_cos, cudaStream_t stream);

// Unweighted chunk mean.
cudaError_t launch_chunk_mean(const float* chunks, int G, int C, int D, float* out, cudaStream_t stream);

}  // namespace qc
