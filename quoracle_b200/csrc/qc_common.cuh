// qc_common.cuh -- launch descriptors shared between the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace qc {

// One launch path for every kernel of the library: `pdl` adds the programmatic-stream-serialization attribute (the
// kernel must then call ptx::griddep_wait() before touching memory of earlier kernels in the stream).
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// K1 tiling variant: tokens per resident block, CTAs per MMA group (1 | 2), k-width of a streamed W stage.
struct EncoderVariant {
  int token_block;
  int cta_group;
  int w_bk;
  int gather;   // 0 = cp.async (LDGSTS) gather, 1 = TMA gather4
  int mcast;    // CTA pairs per cluster sharing each W^T tile through TMA multicast: 1 (none), 2 or 4
};

struct EncoderLaunch {
  const CUtensorMap* w_tmap64; // W^T, box 64 x 128, 128-byte swizzle (host copy; passed __grid_constant__)
  const CUtensorMap* w_tmap32; // W^T, box 32 x 128, 64-byte swizzle
  const CUtensorMap* e_tmap;   // E [vocab x dim], box 64 x 1, 128-byte swizzle (gather4)
  const uint16_t* table;       // device, bf16 [vocab x dim]
  const int32_t* token_ids;    // device [R x T]
  const int32_t* lens;         // device [R]; nullptr = every block full (packed token stream)
  float* part;                 // device [R x ceil(T/TB) x dim]
  int32_t* err_flag;           // device int32[8]: [0..3] watchdog record, [4] token-range flag
  unsigned long long* stall;   // device [grid x 16] per-role stall cycles, or nullptr
  const int32_t* block_map;    // device: ids of the non-empty token blocks (cand * bpc + j), or nullptr = dense
  int nslots;                  // entries in block_map
  int chunked;                 // 1: packed token stream; part rows are addressed through chunk_row
  const int32_t* chunk_row;    // device [R x bpc x (TB/CH)]: partial-row id of every CH-token chunk, -1 for padding
  int R, T, dim, vocab, num_sms;
  int max_clusters;            // co-resident clusters of this variant (0 = ask the occupancy API at launch)
  EncoderVariant variant;
  cudaStream_t stream;
  bool pdl;                    // launch with programmatic stream serialization (see launch_kernel)
};

// Per-device function attributes (dynamic shared-memory opt-ins); qc_ctx_create calls these with the context's
// device current -- there is no process-wide "already set" state anywhere in the library.
cudaError_t encoder_prepare_device(int dim);
cudaError_t similarity_prepare_device();
cudaError_t similarity_tc_prepare_device();

// K1
cudaError_t launch_encode(const EncoderLaunch& a);
bool encoder_variant_supported(int dim, const EncoderVariant& v);
int encoder_variant_w_stages(int dim, const EncoderVariant& v);
EncoderVariant encoder_default_variant(int dim);
EncoderVariant encoder_auto_variant(int dim, int T);

// K2: row r -> L2normalise( (sum_j part[r, j, :]) / len[r] ), j < ceil(len/TB); len == 0 -> zeros.  The row is
// stored to `count` destinations (one for the plain path; this rank's and every peer's gathered buffer when a
// pool is split across GPUs) at row (r / n_local) * n_total + row_offset + r % n_local.
constexpr int kMaxFinalizeDst = 8;
struct FinalizeDst {
  float* base[kMaxFinalizeDst];
  int count;
  int n_local;     // candidates per decision held by this rank
  int n_total;     // candidates per decision in the destination layout
  int row_offset;  // rank * n_local
};
inline FinalizeDst finalize_plain(float* emb) {
  FinalizeDst d{};
  d.base[0] = emb;
  d.count = 1;
  d.n_local = 1;
  d.n_total = 1;
  d.row_offset = 0;
  return d;
}
cudaError_t launch_finalize(const float* part, const int32_t* lens, const FinalizeDst& dst, const int32_t* row_start,
                            int R, int bpc, int TB, int dim, cudaStream_t stream, bool pdl = false);
// Ragged batches: token ids copied into a stream of `chunk`-slot chunks (each chunk belongs to one candidate);
// chunk = encoder_chunk_tokens(token block): the widest of 32 / 16 / 8 dividing it.
cudaError_t launch_pack_tokens(const int32_t* token_ids, const int32_t* chunk_src, const int32_t* chunk_n,
                               int32_t* packed, int nchunks, int chunk, cudaStream_t stream, bool pdl = false);
inline int encoder_chunk_tokens(int tb) { return tb % 32 == 0 ? 32 : (tb % 16 == 0 ? 16 : 8); }

// Candidates sharded on T: per-rank partial sums [world][max_rows][D] + token counts [world][max_rows] in every
// rank's exchange buffer; psum[q] / lens[q] = rank q's copy (peer-mapped).
struct TsplitDst {
  float* psum[kMaxFinalizeDst];
  int32_t* lens[kMaxFinalizeDst];
  int count;     // world
  int rank;
  int max_rows;
};
cudaError_t launch_tsplit_partial(const float* part, const int32_t* lens, const TsplitDst& dst, int R, int bpc, int TB,
                                  int dim, cudaStream_t stream);
cudaError_t launch_tsplit_reduce(const float* psum, const int32_t* lens, float* emb, int R, int world, int max_rows,
                                 int dim, cudaStream_t stream);

// Split-pool handshake: flags[p] = device pointer to rank p's flag array (uint32[kMaxFinalizeDst], peer-mapped).
struct PeerFlags {
  uint32_t* flags[kMaxFinalizeDst];
  int world;
  int rank;
};
cudaError_t launch_peer_signal(const PeerFlags& f, uint32_t epoch, cudaStream_t stream);
cudaError_t launch_peer_wait(const PeerFlags& f, uint32_t epoch, int32_t* dbg, cudaStream_t stream);

// K3: per decision: N x N cosine matrix, reference verdict, medoid vote.
cudaError_t launch_similarity(const float* emb, int B, int N, int D, const float* thr, int32_t* ref_ok,
                              int32_t* medoid, float* cos, int num_sms, cudaStream_t stream, bool pdl = false);

// K3 for pools of 9..64 on tcgen05 (kind::tf32, 3xTF32 split); x_tmap views emb as [B*N x D] fp32, box 32 x 128.
cudaError_t launch_similarity_tc(const CUtensorMap* x_tmap, int B, int N, int D, const float* thr, int32_t* ref_ok,
                                 int32_t* medoid, float* cos, int32_t* dbg, unsigned long long* stall, int num_sms,
                                 cudaStream_t stream, bool pdl = false);

// Lesson dedup: first m with cos(query_q, rows[q, m]) >= thr[q].
cudaError_t launch_first_match(const float* query, const float* rows, int Q, int M, int D, const float* thr,
                               int32_t* out_index, float* out_cos, cudaStream_t stream);

// Unweighted chunk mean.
cudaError_t launch_chunk_mean(const float* chunks, int G, int C, int D, float* out, cudaStream_t stream);

}  // namespace qc
