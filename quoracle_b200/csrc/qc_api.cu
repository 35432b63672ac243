// qc_api.cu -- the C ABI declared in include/qconsensus.h: context, workspaces, host<->device
// staging, kernel sequencing (K1 encode -> K2 finalise -> K3 similarity) and event timing.
// No CPU compute path exists in this library: without a usable sm_100 device every entry
// point returns QC_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/qconsensus.h"
#include "qc_common.cuh"

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes + bytes / 4);
    if (e == cudaSuccess) cap = bytes + bytes / 4;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct ProfRec {
  int kernel;
  cudaEvent_t start, stop;
};

struct PendingOut {
  void* user;
  const void* staged;
  size_t bytes;
};

// One in-flight batch of the pipelined host API (qc_decide_batch_submit / _wait): its own stream,
// device workspaces, pinned staging and status words, so two batches can overlap copy and compute.
struct Slot {
  cudaStream_t stream = nullptr;
  DevBuf tok, lens, thr, emb, refok, medoid, cosm, part, map, packed;
  HostBuf h_in, h_out, h_flags, h_map;
  std::vector<PendingOut> pend;
  bool busy = false;
};

// Workspaces of the device entry points, one set per caller stream (up to four): launches enqueued on different
// streams may overlap on the GPU, so they must not share the K1 -> K2 partial-sum buffer or the embedding scratch.
struct StreamWs {
  cudaStream_t s = nullptr;
  bool used = false;
  unsigned long long tick = 0;
  DevBuf part, emb;
};

// How the NEXT encoder launch of a context walks its tokens.  The host entry points (which hold the lengths) fill it
// right before they call encode_dev, which consumes it exactly once; device callers can set the block list through
// qc_encoder_hint_block_map.  Everything points into device memory that stays valid until that launch has run.
struct EncodePlan {
  // compacted list of the non-empty token blocks (dense walk when null)
  const int32_t* block_map = nullptr;
  int nslots = 0;
  // packed token stream (ragged batches): per chunk its source offset, valid count and partial-row id; per
  // candidate the range of its partial rows
  const int32_t* chunk_src = nullptr;
  const int32_t* chunk_n = nullptr;
  const int32_t* chunk_row = nullptr;
  const int32_t* row_start = nullptr;
  int nchunks = 0;
  int nruns = 0;
  DevBuf* packed_buf = nullptr;   // where the pack kernel writes the stream
  bool variant_set = false;       // the plan was made for ctx->variant: encode_dev must not re-pick the tiling
  void clear() { *this = EncodePlan(); }
};

}  // namespace

struct qc_ctx {
  int device = -1;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  // encoder
  bool has_encoder = false;
  int dim = 0, vocab = 0;
  qc::EncoderVariant variant{0, 0, 0, 0, 1};
  int mcast = 0;                 // W multicast degree forced by qc_encoder_set_multicast (0 = the variant's own)
  EncodePlan plan;               // one-shot plan of the next encoder launch (see EncodePlan)
  DevBuf map, packed;
  HostBuf h_map;
  StreamWs ws[4];
  unsigned long long ws_tick = 0;
  int packing = 0;               // 0 auto, 1 never, 2 whenever the tiling allows (test hook)
  int sim_path = 0;              // K3 for pools 9..64: 0 auto (tcgen05 when the layout allows), 1 mma.sync, 2 tcgen05
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved once at create
  bool pdl = true;               // programmatic dependent launch of K1 / K2 / K3 (QC_PDL=0 in the environment turns it off)
  bool variant_forced = false;   // set by qc_encoder_set_variant / _set_cta_group; else chosen per call from T
  uint16_t* d_table = nullptr;
  uint16_t* d_wt = nullptr;  // W^T [dim(out) x dim(in)] bf16
  CUtensorMap w_tmap64, w_tmap32, e_tmap;
  // Device status words, 8 per block: block 0 for the synchronous and _dev entry points, block 1 + k for pipeline
  // slot k.  Every in-flight batch reads and clears only its own block, so a flag raised by one slot's kernels can
  // neither fail a neighbouring slot's valid batch nor be cleared by it before its owner has looked.
  int32_t* d_err = nullptr;
  int32_t hang[4] = {};
  // workspaces
  DevBuf part, emb, tok, lens, thr, cosm, refok, medoid, q, rows, idx, stall;
  Slot slots[QC_PIPELINE_SLOTS];
  bool stall_on = false;
  HostBuf h_in, h_out;
  // counters / profiling
  uint64_t launches = 0;
  bool profiling = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[QC_K_COUNT] = {};
  uint64_t nl[QC_K_COUNT] = {};
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return QC_OK;
  if (e == cudaErrorMemoryAllocation) return QC_ERR_OUT_OF_MEMORY;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return QC_ERR_NO_DEVICE;
  return QC_ERR_CUDA;
}

#define QC_CUDA(expr)                                  \
  do {                                                 \
    cudaError_t _e = (expr);                           \
    if (_e != cudaSuccess) {                           \
      cudaGetLastError();                              \
      return cuda_status(_e);                          \
    }                                                  \
  } while (0)

cudaEvent_t get_event(qc_ctx* c) {
  if (!c->pool.empty()) {
    cudaEvent_t e = c->pool.back();
    c->pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct KernelScope {
  qc_ctx* c;
  int k;
  cudaStream_t s;
  ProfRec r{};
  KernelScope(qc_ctx* c_, int k_, cudaStream_t s_) : c(c_), k(k_), s(s_) {
    c->launches++;
    if (c->profiling) {
      r.kernel = k;
      r.start = get_event(c);
      r.stop = get_event(c);
      record(r.start);
    }
  }
  ~KernelScope() {
    if (c->profiling) {
      record(r.stop);
      c->recs.push_back(r);
    }
  }
  // While the stream is being captured (qc_capture_begin) the bracket becomes a pair of event-record NODES of the graph
  // (cudaEventRecordExternal): the durations are then taken on the device during a replay, with no host time between
  // the record and the launch it brackets.
  void record(cudaEvent_t e) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive)
      cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
    else
      cudaEventRecord(e, s);
  }
};

void drain_profile(qc_ctx* c) {
  for (ProfRec& r : c->recs) {
    cudaEventSynchronize(r.stop);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.start, r.stop) == cudaSuccess) {
      c->ms[r.kernel] += t;
      c->nl[r.kernel] += 1;
    }
    c->pool.push_back(r.start);
    c->pool.push_back(r.stop);
  }
  c->recs.clear();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int resolve_encode_tiled(qc_ctx* c) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) return QC_ERR_CUDA;
  c->encode_tiled = fn;
  return QC_OK;
}

int make_w_tmap(qc_ctx* c) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(c->encode_tiled);
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(c->dim), static_cast<cuuint64_t>(c->dim)};  // {k, feature}
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(c->dim) * 2};
  const cuuint32_t box64[2] = {64, 128};
  const cuuint32_t box32[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&c->w_tmap64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c->d_wt, gdim, gstride, box64, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return QC_ERR_CUDA;
  r = enc(&c->w_tmap32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c->d_wt, gdim, gstride, box32, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return QC_ERR_CUDA;
  // embedding table for TMA gather4: one 128-byte (64 x bf16) row segment per gathered row
  const cuuint64_t edim[2] = {static_cast<cuuint64_t>(c->dim), static_cast<cuuint64_t>(c->vocab)};
  const cuuint64_t estride[1] = {static_cast<cuuint64_t>(c->dim) * 2};
  const cuuint32_t ebox[2] = {64, 1};
  r = enc(&c->e_tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c->d_table, edim, estride, ebox, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? QC_OK : QC_ERR_CUDA;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Stage `bytes` of caller host memory to device: direct when the caller's buffer is pinned,
// through the context's pinned staging area otherwise.  `stage_off` advances through h_in.
int h2d(qc_ctx* c, void* dst, const void* src, size_t bytes, size_t* stage_off) {
  if (bytes == 0) return QC_OK;
  if (is_pinned(src)) {
    QC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return QC_OK;
  }
  char* st = static_cast<char*>(c->h_in.p) + *stage_off;
  memcpy(st, src, bytes);
  *stage_off += (bytes + 255) & ~size_t(255);
  QC_CUDA(cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, c->stream));
  return QC_OK;
}

int d2h(qc_ctx* c, void* user, const void* dsrc, size_t bytes, size_t* stage_off, std::vector<PendingOut>* pend) {
  if (bytes == 0 || user == nullptr) return QC_OK;
  if (is_pinned(user)) {
    QC_CUDA(cudaMemcpyAsync(user, dsrc, bytes, cudaMemcpyDeviceToHost, c->stream));
    return QC_OK;
  }
  char* st = static_cast<char*>(c->h_out.p) + *stage_off;
  *stage_off += (bytes + 255) & ~size_t(255);
  QC_CUDA(cudaMemcpyAsync(st, dsrc, bytes, cudaMemcpyDeviceToHost, c->stream));
  pend->push_back({user, st, bytes});
  return QC_OK;
}

size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

// Ragged batches: the host entry points hold the lengths, so they list the non-empty token blocks
// (cand * bpc + j, in order) and K1 walks that list instead of skipping: no CTA pair runs half empty.  Batches
// without an empty block need no list (dense enumeration is the same walk).  The list rides to the device on
// the call's own stream through a pinned staging buffer that stays untouched until the call is waited for.
// Packed-stream plan.  Candidate r contributes ceil(len[r] / CH) chunks, laid end to end; chunk c belongs to block
// c / nch.  src[c] = index of the chunk's first token in the caller's token array, cnt[c] = its valid tokens,
// crow[c] = id of the partial row its sum is added to (consecutive chunks of one candidate inside one block share a
// row; -1 for the padding chunks that fill the last block), row[r] .. row[r+1] = the candidate's rows.
// Returns the number of partial rows.
template <class BaseOf>
long long fill_pack_plan(const int32_t* lens, int R, int CH, int nch, long long nchunks, BaseOf base_of, int32_t* src,
                         int32_t* cnt, int32_t* crow, int32_t* row) {
  long long k = 0, runs = 0;
  for (int r = 0; r < R; ++r) {
    row[r] = static_cast<int32_t>(runs);
    const int len = lens[r];
    const long long base = base_of(r);
    for (int o = 0; o < len; o += CH, ++k) {
      src[k] = static_cast<int32_t>(base + o);
      cnt[k] = len - o < CH ? len - o : CH;
      if (o == 0 || k % nch == 0) ++runs;                      // new candidate, or the candidate crosses a block edge
      crow[k] = static_cast<int32_t>(runs - 1);
    }
  }
  row[R] = static_cast<int32_t>(runs);
  for (; k < nchunks; ++k) {
    src[k] = 0;
    cnt[k] = 0;
    crow[k] = -1;
  }
  return runs;
}

constexpr int kErrWords = 8;
constexpr int kErrBlocks = 1 + QC_PIPELINE_SLOTS;
inline int32_t* err_block(qc_ctx* c, int slot /* -1: synchronous and _dev calls */) { return c->d_err + (slot + 1) * kErrWords; }

StreamWs* ws_for(qc_ctx* c, cudaStream_t s) {
  for (StreamWs& w : c->ws)
    if (w.used && w.s == s) {
      w.tick = ++c->ws_tick;
      return &w;
    }
  for (StreamWs& w : c->ws)
    if (!w.used) {
      w.used = true;
      w.s = s;
      w.tick = ++c->ws_tick;
      return &w;
    }
  // all sets taken: recycle the least recently used one once its stream has drained (a stream the caller has
  // destroyed meanwhile reports an error here, which only means there is nothing left to wait for)
  StreamWs* lru = &c->ws[0];
  for (StreamWs& w : c->ws)
    if (w.tick < lru->tick) lru = &w;
  if (cudaStreamSynchronize(lru->s) != cudaSuccess) cudaGetLastError();
  lru->s = s;
  lru->tick = ++c->ws_tick;
  return lru;
}

int stage_block_map(qc_ctx* c, const int32_t* lens, int R, int T, DevBuf* dmap, HostBuf* hmap, DevBuf* dpacked,
                    cudaStream_t s) {
  c->plan.clear();
  // tiling for the unpacked walks: the one chosen for T (least padding of full-length candidates)
  const qc::EncoderVariant va = c->variant_forced ? c->variant : qc::encoder_auto_variant(c->dim, T);
  // tiling for the packed stream: padding no longer depends on T, so the fastest full-block tiling
  const qc::EncoderVariant vp = c->variant_forced ? c->variant : qc::encoder_default_variant(c->dim);
  const int TB = va.token_block;
  const int bpc = (T + TB - 1) / TB;
  const int CH = qc::encoder_chunk_tokens(vp.token_block);
  long long used = 0, chunks = 0;
  for (int r = 0; r < R; ++r) {
    used += (lens[r] + TB - 1) / TB;
    chunks += (lens[r] + CH - 1) / CH;
  }
  // Packed token stream: every candidate is cut into CH-token chunks (CH = 32, or 16 / 8 for token blocks that 32
  // does not divide) and the chunks of all candidates are laid end to end, TB / CH to a block, so the only padding left is inside each candidate's last chunk.
  // Worth its extra small kernel when it saves at least 5 % of the token slots K1 has to process.
  const bool can_pack = static_cast<long long>(R) * T < 0x7fffffffLL && c->packing != 1;
  if (can_pack && chunks > 0) {
    const int nch = vp.token_block / CH;
    const long long pblocks = (chunks + nch - 1) / nch;
    if (c->packing == 2 || pblocks * vp.token_block * 100 <= used * TB * 95) {
      c->variant = vp;
      c->plan.variant_set = true;
      const long long nchunks = pblocks * nch;
      const size_t ints = static_cast<size_t>(3 * nchunks + R + 1);
      QC_CUDA(hmap->ensure(ints * sizeof(int32_t)));
      QC_CUDA(dmap->ensure(ints * sizeof(int32_t)));
      int32_t* src = static_cast<int32_t*>(hmap->p);
      const long long runs = fill_pack_plan(lens, R, CH, nch, nchunks, [T](int r) { return static_cast<long long>(r) * T; },
                                            src, src + nchunks, src + 2 * nchunks, src + 3 * nchunks);
      QC_CUDA(cudaMemcpyAsync(dmap->p, src, ints * sizeof(int32_t), cudaMemcpyHostToDevice, s));
      const int32_t* d = static_cast<const int32_t*>(dmap->p);
      c->plan.chunk_src = d;
      c->plan.chunk_n = d + nchunks;
      c->plan.chunk_row = d + 2 * nchunks;
      c->plan.row_start = d + 3 * nchunks;
      c->plan.nchunks = static_cast<int>(nchunks);
      c->plan.nruns = static_cast<int>(runs);
      c->plan.packed_buf = dpacked;
      return QC_OK;
    }
  }
  c->variant = va;
  c->plan.variant_set = true;
  if (used == static_cast<long long>(R) * bpc || used > 0x7fffffffLL) return QC_OK;
  QC_CUDA(hmap->ensure(static_cast<size_t>(used > 0 ? used : 1) * sizeof(int32_t)));
  QC_CUDA(dmap->ensure(static_cast<size_t>(used > 0 ? used : 1) * sizeof(int32_t)));
  int32_t* h = static_cast<int32_t*>(hmap->p);
  size_t k = 0;
  for (int r = 0; r < R; ++r) {
    const int nb = (lens[r] + TB - 1) / TB;
    for (int j = 0; j < nb; ++j) h[k++] = r * bpc + j;
  }
  if (used > 0) QC_CUDA(cudaMemcpyAsync(dmap->p, h, static_cast<size_t>(used) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  c->plan.block_map = static_cast<const int32_t*>(dmap->p);
  c->plan.nslots = static_cast<int>(used);
  return QC_OK;
}

// Drops an unconsumed hint when a host entry point leaves early, so it can never leak into a later call.
struct HintGuard {
  qc_ctx* c;
  explicit HintGuard(qc_ctx* c_) : c(c_) {}
  ~HintGuard() { c->plan.clear(); }
};

int encode_dev(qc_ctx* c, const int32_t* d_tok, const int32_t* d_lens, int R, int T, float* d_emb, cudaStream_t s,
               DevBuf* part_buf = nullptr, const qc::FinalizeDst* scatter = nullptr,
               const qc::TsplitDst* tsplit = nullptr, int32_t* err = nullptr) {
  DevBuf& part = part_buf ? *part_buf : ws_for(c, s)->part;
  const EncodePlan plan = c->plan;                             // consumed: the next launch starts from a clean plan
  c->plan.clear();
  if (!c->variant_forced && !plan.variant_set) c->variant = qc::encoder_auto_variant(c->dim, T);
  const int TB = c->variant.token_block;
  const int bpc = (T + TB - 1) / TB;
  const int32_t* map = plan.block_map;
  const int nslots = plan.nslots;
  const int32_t* chunk_src = plan.chunk_src;
  const int32_t* chunk_n = plan.chunk_n;
  const int32_t* row_start = plan.row_start;
  const int32_t* chunk_row = plan.chunk_row;
  const int nruns = plan.nruns;
  const int nchunks = plan.nchunks;
  DevBuf* packed = plan.packed_buf;
  const int CH = qc::encoder_chunk_tokens(TB);
  const bool pack = chunk_src != nullptr && nchunks > 0 && tsplit == nullptr && packed != nullptr;
  if (pack) {
    // token ids -> stream of CH-slot chunks; K1 sees nchunks / (TB/CH) full blocks and writes one partial row per chunk
    QC_CUDA(packed->ensure(static_cast<size_t>(nchunks) * CH * sizeof(int32_t)));
    QC_CUDA(part.ensure(static_cast<size_t>(nruns > 0 ? nruns : 1) * c->dim * sizeof(float)));
    c->launches++;
    QC_CUDA(qc::launch_pack_tokens(d_tok, chunk_src, chunk_n, static_cast<int32_t*>(packed->p), nchunks, CH, s, c->pdl));
  } else {
    QC_CUDA(part.ensure(static_cast<size_t>(R) * bpc * c->dim * sizeof(float)));
  }
  {
    KernelScope ks(c, QC_K_ENCODE, s);
    qc::EncoderLaunch a{};
    a.w_tmap64 = &c->w_tmap64;
    a.w_tmap32 = &c->w_tmap32;
    a.e_tmap = &c->e_tmap;
    a.table = c->d_table;
    if (pack) {
      a.token_ids = static_cast<const int32_t*>(packed->p);
      a.lens = nullptr;                          // every block of the packed stream is full
      a.R = nchunks / (TB / CH);
      a.T = TB;
      a.chunked = 1;
      a.chunk_row = chunk_row;
    } else {
      a.token_ids = d_tok;
      a.lens = d_lens;
      a.R = R;
      a.T = T;
      a.block_map = map;
      a.nslots = nslots;
    }
    a.part = static_cast<float*>(part.p);
    a.err_flag = err != nullptr ? err : err_block(c, -1);
    a.stall = c->stall_on ? static_cast<unsigned long long*>(c->stall.p) : nullptr;
    a.dim = c->dim;
    a.vocab = c->vocab;
    a.variant = c->variant;
    if (c->mcast > 0) {                                       // forced multicast degree, where the tiling has it
      qc::EncoderVariant vm = c->variant;
      vm.mcast = c->mcast;
      if (qc::encoder_variant_supported(c->dim, vm)) a.variant = vm;
    }
    a.num_sms = c->num_sms;
    a.stream = s;
    a.pdl = c->pdl;
    QC_CUDA(qc::launch_encode(a));
  }
  {
    KernelScope ks(c, QC_K_FINALIZE, s);
    if (tsplit != nullptr) {
      QC_CUDA(qc::launch_tsplit_partial(static_cast<const float*>(part.p), d_lens, *tsplit, R, bpc, TB, c->dim, s));
      return QC_OK;
    }
    QC_CUDA(qc::launch_finalize(static_cast<const float*>(part.p), d_lens, scatter ? *scatter : qc::finalize_plain(d_emb),
                                pack ? row_start : nullptr, R, bpc, TB, c->dim, s, c->pdl));
  }
  return QC_OK;
}

// Synchronises the context stream and reads the device status words written by K1:
// [0..3] barrier watchdog record (a protocol bug ended the kernel early), [4] token id out of range.
int check_token_flag(qc_ctx* c) {
  int32_t flags[kErrWords] = {};
  QC_CUDA(cudaMemcpyAsync(flags, err_block(c, -1), sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
  QC_CUDA(cudaStreamSynchronize(c->stream));
  if (flags[0] != 0 || flags[4] != 0) {
    QC_CUDA(cudaMemsetAsync(err_block(c, -1), 0, sizeof(flags), c->stream));
    QC_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (flags[0] != 0) {
    for (int i = 0; i < 4; ++i) c->hang[i] = flags[i];
    return QC_ERR_KERNEL_TIMEOUT;
  }
  if (flags[4] != 0) return QC_ERR_TOKEN_RANGE;
  return QC_OK;
}

}  // namespace

extern "C" {

int qc_abi_version(void) { return QC_ABI_VERSION; }

const char* qc_strerror(int status) {
  switch (status) {
    case QC_OK: return "ok";
    case QC_ERR_INVALID_INPUT: return "invalid input";
    case QC_ERR_NOT_CONFIGURED: return "encoder not configured";
    case QC_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case QC_ERR_CUDA: return "CUDA runtime error";
    case QC_ERR_OUT_OF_MEMORY: return "out of device or pinned memory";
    case QC_ERR_CONTEXT_LENGTH: return "candidate longer than QC_MAX_TOKENS";
    case QC_ERR_UNSUPPORTED: return "shape not supported by this ABI version";
    case QC_ERR_TOKEN_RANGE: return "token id outside the encoder vocabulary";
    case QC_ERR_KERNEL_TIMEOUT: return "device barrier watchdog fired (see qc_last_hang)";
    default: return "unknown status";
  }
}

const char* qc_status_atom(int status) {
  switch (status) {
    case QC_OK: return "ok";
    case QC_ERR_INVALID_INPUT: return "invalid_input";
    case QC_ERR_NOT_CONFIGURED: return "not_configured";
    case QC_ERR_NO_DEVICE: return "service_unavailable";
    case QC_ERR_CUDA: return "service_unavailable";
    case QC_ERR_OUT_OF_MEMORY: return "service_unavailable";
    case QC_ERR_CONTEXT_LENGTH: return "context_length_exceeded";
    case QC_ERR_UNSUPPORTED: return "bad_request";
    case QC_ERR_TOKEN_RANGE: return "invalid_input";
    case QC_ERR_KERNEL_TIMEOUT: return "service_unavailable";
    default: return "invalid_response";
  }
}

void* qc_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void qc_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int qc_ctx_create(qc_ctx** out, int device_id, const qc_encoder_cfg* enc) {
  if (out == nullptr) return QC_ERR_INVALID_INPUT;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return QC_ERR_NO_DEVICE;
  }
  if (device_id < 0 || device_id >= ndev) return QC_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  QC_CUDA(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) return QC_ERR_NO_DEVICE;  // kernels are sm_100a only
  if (enc != nullptr) {
    if (enc->table_bf16 == nullptr || enc->w_bf16 == nullptr || enc->vocab <= 0) return QC_ERR_INVALID_INPUT;
    if (qc::encoder_default_variant(enc->dim).token_block == 0) return QC_ERR_UNSUPPORTED;
  }
  qc_ctx* c = new (std::nothrow) qc_ctx();
  if (c == nullptr) return QC_ERR_OUT_OF_MEMORY;
  c->device = device_id;
  c->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("QC_PDL")) c->pdl = e[0] != '0';
  DeviceGuard g(device_id);
  int st = QC_OK;
  do {
    if ((st = cuda_status(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking))) != QC_OK) break;
    if ((st = cuda_status(cudaMalloc(&c->d_err, kErrBlocks * kErrWords * sizeof(int32_t)))) != QC_OK) break;
    if ((st = cuda_status(cudaMemset(c->d_err, 0, kErrBlocks * kErrWords * sizeof(int32_t)))) != QC_OK) break;
    if ((st = resolve_encode_tiled(c)) != QC_OK) break;
    // per-device function attributes (shared-memory opt-ins) of every kernel this context can launch
    if ((st = cuda_status(qc::similarity_prepare_device())) != QC_OK) break;
    if ((st = cuda_status(qc::similarity_tc_prepare_device())) != QC_OK) break;
    if (enc != nullptr && (st = cuda_status(qc::encoder_prepare_device(enc->dim))) != QC_OK) break;
    if (enc != nullptr) {
      c->dim = enc->dim;
      c->vocab = enc->vocab;
      const size_t tb = static_cast<size_t>(enc->vocab) * enc->dim * 2;
      const size_t wb = static_cast<size_t>(enc->dim) * enc->dim * 2;
      if ((st = cuda_status(cudaMalloc(&c->d_table, tb))) != QC_OK) break;
      if ((st = cuda_status(cudaMalloc(&c->d_wt, wb))) != QC_OK) break;
      if ((st = cuda_status(cudaMemcpy(c->d_table, enc->table_bf16, tb, cudaMemcpyHostToDevice))) != QC_OK) break;
      std::vector<uint16_t> wt(static_cast<size_t>(enc->dim) * enc->dim);
      for (int k = 0; k < enc->dim; ++k)
        for (int d = 0; d < enc->dim; ++d) wt[static_cast<size_t>(d) * enc->dim + k] = enc->w_bf16[static_cast<size_t>(k) * enc->dim + d];
      if ((st = cuda_status(cudaMemcpy(c->d_wt, wt.data(), wb, cudaMemcpyHostToDevice))) != QC_OK) break;
      if ((st = make_w_tmap(c)) != QC_OK) break;
      c->has_encoder = true;
      c->variant = qc::encoder_default_variant(c->dim);
    }
  } while (false);
  if (st != QC_OK) {
    qc_ctx_destroy(c);
    return st;
  }
  *out = c;
  return QC_OK;
}

void qc_ctx_destroy(qc_ctx* c) {
  if (c == nullptr) return;
  DeviceGuard g(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drain_profile(c);
  for (cudaEvent_t e : c->pool) cudaEventDestroy(e);
  c->part.release();
  c->emb.release();
  c->tok.release();
  c->lens.release();
  c->thr.release();
  c->cosm.release();
  c->refok.release();
  c->medoid.release();
  c->q.release();
  c->rows.release();
  c->idx.release();
  c->stall.release();
  c->map.release();
  for (StreamWs& w : c->ws) {
    w.part.release();
    w.emb.release();
  }
  c->packed.release();
  c->h_map.release();
  for (Slot& sl : c->slots) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    sl.tok.release(); sl.lens.release(); sl.thr.release(); sl.emb.release(); sl.refok.release(); sl.medoid.release();
    sl.cosm.release(); sl.part.release(); sl.h_in.release(); sl.h_out.release(); sl.h_flags.release(); sl.map.release(); sl.h_map.release(); sl.packed.release();
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  c->h_in.release();
  c->h_out.release();
  if (c->d_table) cudaFree(c->d_table);
  if (c->d_wt) cudaFree(c->d_wt);
  if (c->d_err) cudaFree(c->d_err);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int qc_ctx_device(const qc_ctx* c) { return c ? c->device : -1; }
int qc_ctx_dim(const qc_ctx* c) { return (c && c->has_encoder) ? c->dim : 0; }

int qc_ctx_synchronize(qc_ctx* c) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  QC_CUDA(cudaStreamSynchronize(c->stream));
  return QC_OK;
}

uint64_t qc_launch_count(const qc_ctx* c) { return c ? c->launches : 0; }

int qc_check_status(qc_ctx* c) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  return check_token_flag(c);
}

int qc_last_hang(const qc_ctx* c, int32_t info[4]) {
  if (c == nullptr || info == nullptr) return QC_ERR_INVALID_INPUT;
  for (int i = 0; i < 4; ++i) info[i] = c->hang[i];
  return QC_OK;
}

int qc_profile_enable(qc_ctx* c, int on) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  if (!on) drain_profile(c);
  c->profiling = on != 0;
  return QC_OK;
}

int qc_profile_reset(qc_ctx* c) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  drain_profile(c);
  for (int k = 0; k < QC_K_COUNT; ++k) {
    c->ms[k] = 0.0;
    c->nl[k] = 0;
  }
  c->launches = 0;
  return QC_OK;
}

int qc_kernel_stats(qc_ctx* c, double ms[QC_K_COUNT], uint64_t launches[QC_K_COUNT]) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  drain_profile(c);
  for (int k = 0; k < QC_K_COUNT; ++k) {
    if (ms) ms[k] = c->ms[k];
    if (launches) launches[k] = c->nl[k];
  }
  return QC_OK;
}

int qc_encoder_info(const qc_ctx* c, int32_t* cta_group, int32_t* token_block, int32_t* w_stages) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (cta_group) *cta_group = c->variant.cta_group;
  if (token_block) *token_block = c->variant.token_block;
  if (w_stages) *w_stages = qc::encoder_variant_w_stages(c->dim, c->variant);
  return QC_OK;
}

int qc_encoder_multicast(const qc_ctx* c, int32_t* pairs) {
  if (c == nullptr || pairs == nullptr) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  qc::EncoderVariant vm = c->variant;
  if (c->mcast > 0) vm.mcast = c->mcast;
  *pairs = qc::encoder_variant_supported(c->dim, vm) ? (vm.mcast < 1 ? 1 : vm.mcast) : (c->variant.mcast < 1 ? 1 : c->variant.mcast);
  return QC_OK;
}

int qc_encoder_stall_profile(qc_ctx* c, int enable, double out[16]) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  const size_t n = static_cast<size_t>(c->num_sms) * 16;
  QC_CUDA(cudaDeviceSynchronize());
  if (out != nullptr && c->stall.p != nullptr) {
    std::vector<unsigned long long> h(n);
    QC_CUDA(cudaMemcpy(h.data(), c->stall.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    int cnt[16] = {};
    for (int k = 0; k < 16; ++k) out[k] = 0.0;
    for (int b = 0; b < c->num_sms; ++b) {
      for (int k = 0; k < 16; ++k) {
        const int total_slot = k < 4 ? 0 : (k < 6 ? 4 : (k < 8 ? 6 : (k < 10 ? 8 : (k == 10 ? 6 : (k < 13 ? 0 : (k == 13 ? 6 : k))))));
        if (h[b * 16 + total_slot] == 0) continue;   // role did not run in this CTA (e.g. MMA in the peer CTA)
        out[k] += static_cast<double>(h[b * 16 + k]);
        cnt[k]++;
      }
    }
    for (int k = 0; k < 16; ++k) out[k] = cnt[k] ? out[k] / cnt[k] : 0.0;
  }
  if (enable) {
    QC_CUDA(c->stall.ensure(n * sizeof(unsigned long long)));
    QC_CUDA(cudaMemset(c->stall.p, 0, n * sizeof(unsigned long long)));
  }
  c->stall_on = enable != 0;
  return QC_OK;
}

int qc_encoder_set_cta_group(qc_ctx* c, int cta_group) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  // keep the default tiling of this dim when it exists for the requested group size, else the
  // baseline cp.async tiling (every dim has one for cta_group 1; 768/1536 also for 2)
  qc::EncoderVariant v = c->variant;
  v.cta_group = cta_group;
  if (!qc::encoder_variant_supported(c->dim, v)) {
    v = qc::EncoderVariant{c->dim == 1536 ? 64 : 128, cta_group, 64, 0, 1};
    if (!qc::encoder_variant_supported(c->dim, v)) return QC_ERR_UNSUPPORTED;
  }
  c->variant = v;
  c->variant_forced = true;
  return QC_OK;
}

int qc_encoder_set_variant(qc_ctx* c, int token_block, int cta_group, int w_bk, int gather_mode) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (token_block == 0) {             // back to automatic selection
    c->variant_forced = false;
    c->variant = qc::encoder_default_variant(c->dim);
    return QC_OK;
  }
  const qc::EncoderVariant v{token_block, cta_group, w_bk, gather_mode, 1};
  if (!qc::encoder_variant_supported(c->dim, v)) return QC_ERR_UNSUPPORTED;
  c->variant = v;
  c->variant_forced = true;
  return QC_OK;
}

int qc_encoder_set_multicast(qc_ctx* c, int pairs) {
  if (c == nullptr || (pairs != 0 && pairs != 1 && pairs != 2 && pairs != 4)) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  c->mcast = pairs;
  return QC_OK;
}

// ---- device entry points ---------------------------------------------------------------

int qc_embed_batch_dev(qc_ctx* c, const int32_t* d_tok, const int32_t* d_lens, int R, int T, float* d_emb,
                       void* stream) {
  if (c == nullptr || d_tok == nullptr || d_lens == nullptr || d_emb == nullptr || R < 0 || T < 1)
    return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (T > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  if (R == 0) return QC_OK;
  DeviceGuard g(c->device);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  return encode_dev(c, d_tok, d_lens, R, T, d_emb, s);
}

}  // extern "C"

namespace {
// K3 on a caller-chosen status block (the public _dev entry point uses block 0, pipeline slots their own).
int similarity_dev(qc_ctx* c, const float* d_emb, int B, int N, int D, const float* d_thr, int32_t* d_ref_ok,
                   int32_t* d_medoid, float* d_cos, void* stream, int32_t* err) {
  if (c == nullptr || d_emb == nullptr || d_thr == nullptr || B < 0 || D < 1) return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (B == 0) return QC_OK;
  DeviceGuard g(c->device);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  KernelScope ks(c, QC_K_SIMILARITY, s);
  // Pools of 9..64 run on the tensor cores (tcgen05 kind::tf32, 3xTF32 split) when TMA can address the
  // matrix: 16-byte aligned base and row pitch, row count within a 32-bit coordinate.
  const bool tc_ok = N > 8 && N <= 64 && (D % 4) == 0 && (reinterpret_cast<uintptr_t>(d_emb) & 15) == 0 &&
                     static_cast<long long>(B) * N < (1ll << 31) && c->encode_tiled != nullptr;
  if (c->sim_path == 2 && !tc_ok) return QC_ERR_UNSUPPORTED;
  if (tc_ok && c->sim_path != 1) {
    CUtensorMap xm;
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(B) * N};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(D) * 4};
    // a tile is floor(128 / N) whole decisions: only their rows are fetched (the rest of the 128-row MMA tile
    // keeps stale shared memory, which only reaches accumulator entries that are discarded)
    const cuuint32_t box[2] = {32, static_cast<cuuint32_t>((128 / N) * N)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(c->encode_tiled)(
        &xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d_emb), gdim, gstride, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return QC_ERR_CUDA;
    QC_CUDA(qc::launch_similarity_tc(&xm, B, N, D, d_thr, d_ref_ok, d_medoid, d_cos, err,
                                     c->stall_on ? static_cast<unsigned long long*>(c->stall.p) : nullptr, c->num_sms, s, c->pdl));
    return QC_OK;
  }
  QC_CUDA(qc::launch_similarity(d_emb, B, N, D, d_thr, d_ref_ok, d_medoid, d_cos, c->num_sms, s, c->pdl));
  return QC_OK;
}
}  // namespace

extern "C" {

int qc_similarity_batch_dev(qc_ctx* c, const float* d_emb, int B, int N, int D, const float* d_thr,
                            int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos, void* stream) {
  if (c == nullptr) return QC_ERR_INVALID_INPUT;
  return similarity_dev(c, d_emb, B, N, D, d_thr, d_ref_ok, d_medoid, d_cos, stream, err_block(c, -1));
}

// Host-only: the packed-stream plan exactly as the host entry points build it (no device, no context needed), so the
// planner that drives K1's gather and K2's row ranges can be checked where there is no GPU.
int qc_debug_pack_plan(const int32_t* lens, int R, int T, int token_block, int32_t* out, int64_t out_capacity,
                       int32_t* nchunks_out, int32_t* nruns_out) {
  if (lens == nullptr || out == nullptr || nchunks_out == nullptr || nruns_out == nullptr || R < 0 || T < 1 ||
      token_block < 8 || token_block % 8 != 0)
    return QC_ERR_INVALID_INPUT;
  const int CH = qc::encoder_chunk_tokens(token_block), nch = token_block / CH;
  long long chunks = 0;
  for (int r = 0; r < R; ++r) {
    if (lens[r] < 0 || lens[r] > T) return QC_ERR_INVALID_INPUT;
    chunks += (lens[r] + CH - 1) / CH;
  }
  const long long nchunks = (chunks + nch - 1) / nch * nch;
  if (3 * nchunks + R + 1 > out_capacity) return QC_ERR_INVALID_INPUT;
  const long long runs = fill_pack_plan(lens, R, CH, nch, nchunks, [T](int r) { return static_cast<long long>(r) * T; }, out,
                                        out + nchunks, out + 2 * nchunks, out + 3 * nchunks);
  *nchunks_out = static_cast<int32_t>(nchunks);
  *nruns_out = static_cast<int32_t>(runs);
  return QC_OK;
}

int qc_encoder_set_packing(qc_ctx* c, int mode) {
  if (c == nullptr || mode < 0 || mode > 2) return QC_ERR_INVALID_INPUT;
  c->packing = mode;
  return QC_OK;
}

int qc_encoder_hint_block_map(qc_ctx* c, const int32_t* d_block_map, int nslots) {
  if (c == nullptr || nslots < 0 || (d_block_map == nullptr && nslots != 0)) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  c->plan.block_map = d_block_map;
  c->plan.nslots = d_block_map ? nslots : 0;
  return QC_OK;
}

int qc_encoder_token_block(qc_ctx* c, int T, int32_t* token_block) {
  if (c == nullptr || token_block == nullptr || T < 1) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  *token_block = (c->variant_forced ? c->variant : qc::encoder_auto_variant(c->dim, T)).token_block;
  return QC_OK;
}

int qc_similarity_set_path(qc_ctx* c, int path) {
  if (c == nullptr || path < 0 || path > 2) return QC_ERR_INVALID_INPUT;
  c->sim_path = path;
  return QC_OK;
}

int qc_decide_batch_dev(qc_ctx* c, const int32_t* d_tok, const int32_t* d_lens, int B, int N, int T,
                        const float* d_thr, int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos, float* d_emb,
                        void* stream) {
  if (c == nullptr || d_tok == nullptr || d_lens == nullptr || d_thr == nullptr || B < 0 || T < 1)
    return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (T > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  if (B == 0) return QC_OK;
  DeviceGuard g(c->device);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  const int R = B * N;
  float* emb = d_emb;
  if (emb == nullptr) {
    DevBuf& scratch = ws_for(c, s)->emb;
    QC_CUDA(scratch.ensure(static_cast<size_t>(R) * c->dim * sizeof(float)));
    emb = static_cast<float*>(scratch.p);
  }
  int st = encode_dev(c, d_tok, d_lens, R, T, emb, s);
  if (st != QC_OK) return st;
  return qc_similarity_batch_dev(c, emb, B, N, c->dim, d_thr, d_ref_ok, d_medoid, d_cos, s);
}

// ---- host entry points -------------------------------------------------------------------

int qc_embed_batch(qc_ctx* c, const int32_t* token_ids, const int32_t* lens, int R, int T, float* out_emb) {
  if (c == nullptr || token_ids == nullptr || lens == nullptr || out_emb == nullptr || R < 0 || T < 1)
    return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (T > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  if (R == 0) return QC_OK;
  for (int r = 0; r < R; ++r)
    if (lens[r] < 0 || lens[r] > T) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  const size_t tokb = static_cast<size_t>(R) * T * 4, lenb = static_cast<size_t>(R) * 4;
  const size_t embb = static_cast<size_t>(R) * c->dim * 4;
  QC_CUDA(c->tok.ensure(tokb));
  QC_CUDA(c->lens.ensure(lenb));
  QC_CUDA(c->emb.ensure(embb));
  QC_CUDA(c->h_in.ensure(pad256(tokb) + pad256(lenb)));
  QC_CUDA(c->h_out.ensure(pad256(embb)));
  size_t off = 0;
  int st;
  if ((st = h2d(c, c->tok.p, token_ids, tokb, &off)) != QC_OK) return st;
  if ((st = h2d(c, c->lens.p, lens, lenb, &off)) != QC_OK) return st;
  HintGuard hg(c);
  if ((st = stage_block_map(c, lens, R, T, &c->map, &c->h_map, &c->packed, c->stream)) != QC_OK) return st;
  st = encode_dev(c, static_cast<int32_t*>(c->tok.p), static_cast<int32_t*>(c->lens.p), R, T,
                  static_cast<float*>(c->emb.p), c->stream);
  if (st != QC_OK) return st;
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_emb, c->emb.p, embb, &ooff, &pend)) != QC_OK) return st;
  if ((st = check_token_flag(c)) != QC_OK) return st;
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

// ---- ragged (CSR) input: no padding on the host, none on the wire -------------------------------
// tokens of candidate r are tokens[offsets[r] .. offsets[r+1]).  The flat array goes to the device as is; the packed
// token stream (chunks of every candidate laid end to end) is built from it by the pack kernel, so the only padding
// K1 ever sees is inside each candidate's last chunk.
struct RaggedBufs {           // the context's own buffers (synchronous calls) or one pipeline slot's
  DevBuf* tok;
  DevBuf* lens;
  DevBuf* map;
  DevBuf* packed;
  DevBuf* part;
  HostBuf* h_in;
  HostBuf* h_map;
};

static int ragged_common(qc_ctx* c, const RaggedBufs& wb, const int32_t* tokens, const int32_t* offsets, int R,
                         float* d_emb, cudaStream_t s, int32_t* err = nullptr) {
  const long long total = offsets[R];
  const size_t tokb = static_cast<size_t>(total > 0 ? total : 1) * 4, lenb = static_cast<size_t>(R) * 4;
  QC_CUDA(wb.tok->ensure(tokb));
  QC_CUDA(wb.lens->ensure(lenb));
  QC_CUDA(wb.h_in->ensure(pad256(tokb) + pad256(lenb)));
  // lengths are built straight into the pinned staging area (it stays untouched until the call is waited for)
  int32_t* lens = reinterpret_cast<int32_t*>(static_cast<char*>(wb.h_in->p) + pad256(tokb));
  for (int r = 0; r < R; ++r) {
    const long long l = static_cast<long long>(offsets[r + 1]) - offsets[r];
    if (offsets[r] < 0 || l < 0) return QC_ERR_INVALID_INPUT;
    if (l > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
    lens[r] = static_cast<int32_t>(l);
  }
  if (total > 0) {
    const void* src = tokens;
    if (!is_pinned(tokens)) {
      memcpy(wb.h_in->p, tokens, static_cast<size_t>(total) * 4);
      src = wb.h_in->p;
    }
    QC_CUDA(cudaMemcpyAsync(wb.tok->p, src, static_cast<size_t>(total) * 4, cudaMemcpyHostToDevice, s));
  }
  QC_CUDA(cudaMemcpyAsync(wb.lens->p, lens, lenb, cudaMemcpyHostToDevice, s));
  if (total == 0) {                                            // nothing to encode: zero rows
    QC_CUDA(cudaMemsetAsync(d_emb, 0, static_cast<size_t>(R) * c->dim * 4, s));
    return QC_OK;
  }
  // the plan of stage_block_map's packed branch, with flat offsets as chunk sources
  const qc::EncoderVariant vp = c->variant_forced ? c->variant : qc::encoder_default_variant(c->dim);
  const int CH = qc::encoder_chunk_tokens(vp.token_block), nch = vp.token_block / CH;
  long long chunks = 0;
  for (int r = 0; r < R; ++r) chunks += (lens[r] + CH - 1) / CH;
  const long long nchunks = (chunks + nch - 1) / nch * nch;
  if (nchunks * CH > 0x7fffffffLL) return QC_ERR_UNSUPPORTED;
  const size_t ints = static_cast<size_t>(3 * nchunks + R + 1);
  QC_CUDA(wb.h_map->ensure(ints * sizeof(int32_t)));
  QC_CUDA(wb.map->ensure(ints * sizeof(int32_t)));
  int32_t* src = static_cast<int32_t*>(wb.h_map->p);
  const long long runs = fill_pack_plan(lens, R, CH, nch, nchunks, [offsets](int r) { return static_cast<long long>(offsets[r]); },
                                        src, src + nchunks, src + 2 * nchunks, src + 3 * nchunks);
  QC_CUDA(cudaMemcpyAsync(wb.map->p, src, ints * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  HintGuard hg(c);
  const int32_t* d = static_cast<const int32_t*>(wb.map->p);
  c->variant = vp;
  c->plan.variant_set = true;
  c->plan.chunk_src = d;
  c->plan.chunk_n = d + nchunks;
  c->plan.chunk_row = d + 2 * nchunks;
  c->plan.row_start = d + 3 * nchunks;
  c->plan.nchunks = static_cast<int>(nchunks);
  c->plan.nruns = static_cast<int>(runs);
  c->plan.packed_buf = wb.packed;
  // T only sizes the dense walk, which the pack hint replaces
  return encode_dev(c, static_cast<int32_t*>(wb.tok->p), static_cast<int32_t*>(wb.lens->p), R, vp.token_block, d_emb, s,
                    wb.part, nullptr, nullptr, err);
}

static RaggedBufs ctx_ragged_bufs(qc_ctx* c) {
  return RaggedBufs{&c->tok, &c->lens, &c->map, &c->packed, nullptr, &c->h_in, &c->h_map};
}

int qc_embed_ragged(qc_ctx* c, const int32_t* tokens, const int32_t* offsets, int R, float* out_emb) {
  if (c == nullptr || offsets == nullptr || out_emb == nullptr || R < 0) return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (R == 0) return QC_OK;
  if (offsets[R] < 0 || (offsets[R] > 0 && tokens == nullptr)) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  const size_t embb = static_cast<size_t>(R) * c->dim * 4;
  QC_CUDA(c->emb.ensure(embb));
  QC_CUDA(c->h_out.ensure(pad256(embb)));
  int st = ragged_common(c, ctx_ragged_bufs(c), tokens, offsets, R, static_cast<float*>(c->emb.p), c->stream);
  if (st != QC_OK) return st;
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_emb, c->emb.p, embb, &ooff, &pend)) != QC_OK) return st;
  if ((st = check_token_flag(c)) != QC_OK) return st;
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

int qc_decide_ragged(qc_ctx* c, const int32_t* tokens, const int32_t* offsets, int B, int N, const float* thr,
                     int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos, float* out_emb) {
  if (c == nullptr || offsets == nullptr || thr == nullptr || B < 0) return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (B == 0) return QC_OK;
  const int R = B * N;
  if (offsets[R] < 0 || (offsets[R] > 0 && tokens == nullptr)) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  const size_t thrb = static_cast<size_t>(B) * 4, ib = static_cast<size_t>(B) * 4;
  const size_t embb = static_cast<size_t>(R) * c->dim * 4, cosb = static_cast<size_t>(B) * N * N * 4;
  QC_CUDA(c->thr.ensure(thrb));
  QC_CUDA(c->emb.ensure(embb));
  QC_CUDA(c->refok.ensure(ib));
  QC_CUDA(c->medoid.ensure(ib));
  if (out_cos) QC_CUDA(c->cosm.ensure(cosb));
  QC_CUDA(c->h_out.ensure(2 * pad256(ib) + pad256(cosb) + pad256(embb)));
  int st = ragged_common(c, ctx_ragged_bufs(c), tokens, offsets, R, static_cast<float*>(c->emb.p), c->stream);
  if (st != QC_OK) return st;
  // thr rides after the tokens: stage it through its own small pinned copy on the same stream
  QC_CUDA(cudaMemcpyAsync(c->thr.p, thr, thrb, cudaMemcpyHostToDevice, c->stream));
  st = qc_similarity_batch_dev(c, static_cast<float*>(c->emb.p), B, N, c->dim, static_cast<float*>(c->thr.p),
                               static_cast<int32_t*>(c->refok.p), static_cast<int32_t*>(c->medoid.p),
                               out_cos ? static_cast<float*>(c->cosm.p) : nullptr, c->stream);
  if (st != QC_OK) return st;
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_ref_ok, c->refok.p, ib, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_medoid, c->medoid.p, ib, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_cos, c->cosm.p, cosb, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_emb, c->emb.p, embb, &ooff, &pend)) != QC_OK) return st;
  if ((st = check_token_flag(c)) != QC_OK) return st;
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

// ---- pool split across GPUs -------------------------------------------
// Each rank embeds its pool_local candidates of every decision; K2 stores each normalised row into the
// gathered [B, world * pool_local, D] buffer of every rank (its own and, through CUDA IPC peer mappings over
// NVLink, each peer's) -- the all-gather is those stores -- then a flag handshake and K3 on the local copy.
// Two buffers alternate by call parity: a rank can only be two calls ahead of a peer that has not yet read.
struct qc_split {
  qc_ctx* ctx = nullptr;
  int world = 0, rank = 0, max_decisions = 0, pool_local = 0;
  size_t buf_bytes = 0;                       // one gathered buffer
  char* local = nullptr;                      // [2 x buf_bytes][flags]
  char* peer[qc::kMaxFinalizeDst] = {};       // peer-mapped bases (peer[rank] == local)
  bool attached[qc::kMaxFinalizeDst] = {};
  uint32_t epoch = 0;
  int last_buf = 0, last_B = 0;
  int mode = 0;                               // 0: pool split (rows gathered), 1: candidates sharded on T
  int max_rows = 0;                           // mode 1
  size_t psum_bytes = 0;                      // mode 1: [world][max_rows][D] fp32, followed by the token counts
};

int qc_split_create(qc_ctx* c, int world, int rank, int max_decisions, int pool_local, qc_split** out) {
  if (c == nullptr || out == nullptr || world < 1 || world > qc::kMaxFinalizeDst || rank < 0 || rank >= world ||
      max_decisions < 1 || pool_local < 1)
    return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (world * pool_local > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  DeviceGuard g(c->device);
  qc_split* sp = new (std::nothrow) qc_split();
  if (sp == nullptr) return QC_ERR_OUT_OF_MEMORY;
  sp->ctx = c;
  sp->world = world;
  sp->rank = rank;
  sp->max_decisions = max_decisions;
  sp->pool_local = pool_local;
  sp->buf_bytes = pad256(static_cast<size_t>(max_decisions) * world * pool_local * c->dim * sizeof(float));
  const size_t total = 2 * sp->buf_bytes + 256;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, total);      // cudaMalloc: the allocation kind CUDA IPC can export
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete sp;
    return cuda_status(e);
  }
  cudaMemset(p, 0, total);
  cudaDeviceSynchronize();
  sp->local = static_cast<char*>(p);
  sp->peer[rank] = sp->local;
  sp->attached[rank] = true;
  *out = sp;
  return QC_OK;
}

int qc_split_export(qc_split* sp, uint8_t handle[QC_SPLIT_HANDLE_BYTES]) {
  if (sp == nullptr || handle == nullptr) return QC_ERR_INVALID_INPUT;
  static_assert(sizeof(cudaIpcMemHandle_t) <= QC_SPLIT_HANDLE_BYTES, "IPC handle does not fit the ABI handle");
  DeviceGuard g(sp->ctx->device);
  cudaIpcMemHandle_t h;
  QC_CUDA(cudaIpcGetMemHandle(&h, sp->local));
  memset(handle, 0, QC_SPLIT_HANDLE_BYTES);
  memcpy(handle, &h, sizeof(h));
  return QC_OK;
}

int qc_split_attach(qc_split* sp, int peer_rank, const uint8_t handle[QC_SPLIT_HANDLE_BYTES]) {
  if (sp == nullptr || handle == nullptr || peer_rank < 0 || peer_rank >= sp->world) return QC_ERR_INVALID_INPUT;
  if (peer_rank == sp->rank || sp->attached[peer_rank]) return QC_OK;
  DeviceGuard g(sp->ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  QC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  sp->peer[peer_rank] = static_cast<char*>(p);
  sp->attached[peer_rank] = true;
  return QC_OK;
}

int qc_split_destroy(qc_split* sp) {
  if (sp == nullptr) return QC_OK;
  DeviceGuard g(sp->ctx->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < sp->world; ++p)
    if (p != sp->rank && sp->attached[p] && sp->peer[p]) cudaIpcCloseMemHandle(sp->peer[p]);
  if (sp->local) cudaFree(sp->local);
  delete sp;
  return QC_OK;
}

int qc_split_gathered(const qc_split* sp, const float** d_emb, int32_t* decisions) {
  if (sp == nullptr || d_emb == nullptr || sp->mode != 0) return QC_ERR_INVALID_INPUT;
  *d_emb = reinterpret_cast<const float*>(sp->local + static_cast<size_t>(sp->last_buf) * sp->buf_bytes);
  if (decisions) *decisions = sp->last_B;
  return QC_OK;
}

int qc_decide_split_dev(qc_ctx* c, qc_split* sp, const int32_t* d_tok, const int32_t* d_lens, int B, int T,
                        const float* d_thr, int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos, void* stream) {
  if (c == nullptr || sp == nullptr || sp->ctx != c || sp->mode != 0 || d_tok == nullptr || d_lens == nullptr ||
      d_thr == nullptr || B < 0 || T < 1)
    return QC_ERR_INVALID_INPUT;
  if (B > sp->max_decisions) return QC_ERR_UNSUPPORTED;
  if (T > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  for (int p = 0; p < sp->world; ++p)
    if (!sp->attached[p]) return QC_ERR_NOT_CONFIGURED;
  if (B == 0) return QC_OK;
  DeviceGuard g(c->device);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  const uint32_t epoch = ++sp->epoch;
  const int buf = static_cast<int>(epoch & 1u);
  const int n_total = sp->world * sp->pool_local;
  qc::FinalizeDst dst{};
  qc::PeerFlags pf{};
  dst.count = sp->world;
  dst.n_local = sp->pool_local;
  dst.n_total = n_total;
  dst.row_offset = sp->rank * sp->pool_local;
  pf.world = sp->world;
  pf.rank = sp->rank;
  for (int p = 0; p < sp->world; ++p) {
    dst.base[p] = reinterpret_cast<float*>(sp->peer[p] + static_cast<size_t>(buf) * sp->buf_bytes);
    pf.flags[p] = reinterpret_cast<uint32_t*>(sp->peer[p] + 2 * sp->buf_bytes);
  }
  int st = encode_dev(c, d_tok, d_lens, B * sp->pool_local, T, nullptr, s, nullptr, &dst);
  if (st != QC_OK) return st;
  c->launches += 2;
  QC_CUDA(qc::launch_peer_signal(pf, epoch, s));
  QC_CUDA(qc::launch_peer_wait(pf, epoch, err_block(c, -1), s));
  sp->last_buf = buf;
  sp->last_B = B;
  return qc_similarity_batch_dev(c, reinterpret_cast<const float*>(sp->local + static_cast<size_t>(buf) * sp->buf_bytes), B,
                                 n_total, c->dim, d_thr, d_ref_ok, d_medoid, d_cos, s);
}

// ---- candidates sharded on T across GPUs (long-context path) -------------
int qc_tsplit_create(qc_ctx* c, int world, int rank, int max_rows, qc_split** out) {
  if (c == nullptr || out == nullptr || world < 1 || world > qc::kMaxFinalizeDst || rank < 0 || rank >= world ||
      max_rows < 1)
    return QC_ERR_INVALID_INPUT;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  DeviceGuard g(c->device);
  qc_split* sp = new (std::nothrow) qc_split();
  if (sp == nullptr) return QC_ERR_OUT_OF_MEMORY;
  sp->ctx = c;
  sp->world = world;
  sp->rank = rank;
  sp->mode = 1;
  sp->max_rows = max_rows;
  sp->psum_bytes = pad256(static_cast<size_t>(world) * max_rows * c->dim * sizeof(float));
  sp->buf_bytes = sp->psum_bytes + pad256(static_cast<size_t>(world) * max_rows * sizeof(int32_t));
  const size_t total = 2 * sp->buf_bytes + 256;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, total);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete sp;
    return cuda_status(e);
  }
  cudaMemset(p, 0, total);
  cudaDeviceSynchronize();
  sp->local = static_cast<char*>(p);
  sp->peer[rank] = sp->local;
  sp->attached[rank] = true;
  *out = sp;
  return QC_OK;
}

int qc_embed_tsplit_dev(qc_ctx* c, qc_split* sp, const int32_t* d_tok, const int32_t* d_lens, int R, int T_local,
                        float* d_emb, void* stream) {
  if (c == nullptr || sp == nullptr || sp->ctx != c || sp->mode != 1 || d_tok == nullptr || d_lens == nullptr ||
      d_emb == nullptr || R < 0 || T_local < 1)
    return QC_ERR_INVALID_INPUT;
  if (R > sp->max_rows) return QC_ERR_UNSUPPORTED;
  if (T_local > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  for (int p = 0; p < sp->world; ++p)
    if (!sp->attached[p]) return QC_ERR_NOT_CONFIGURED;
  if (R == 0) return QC_OK;
  DeviceGuard g(c->device);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  const uint32_t epoch = ++sp->epoch;
  const size_t off = static_cast<size_t>(epoch & 1u) * sp->buf_bytes;
  qc::TsplitDst dst{};
  qc::PeerFlags pf{};
  dst.count = sp->world;
  dst.rank = sp->rank;
  dst.max_rows = sp->max_rows;
  pf.world = sp->world;
  pf.rank = sp->rank;
  for (int p = 0; p < sp->world; ++p) {
    dst.psum[p] = reinterpret_cast<float*>(sp->peer[p] + off);
    dst.lens[p] = reinterpret_cast<int32_t*>(sp->peer[p] + off + sp->psum_bytes);
    pf.flags[p] = reinterpret_cast<uint32_t*>(sp->peer[p] + 2 * sp->buf_bytes);
  }
  int st = encode_dev(c, d_tok, d_lens, R, T_local, nullptr, s, nullptr, nullptr, &dst);
  if (st != QC_OK) return st;
  c->launches += 3;
  QC_CUDA(qc::launch_peer_signal(pf, epoch, s));
  QC_CUDA(qc::launch_peer_wait(pf, epoch, err_block(c, -1), s));
  QC_CUDA(qc::launch_tsplit_reduce(reinterpret_cast<const float*>(sp->local + off),
                                   reinterpret_cast<const int32_t*>(sp->local + off + sp->psum_bytes), d_emb, R, sp->world,
                                   sp->max_rows, c->dim, s));
  return QC_OK;
}

int qc_decide_tsplit_dev(qc_ctx* c, qc_split* sp, const int32_t* d_tok, const int32_t* d_lens, int B, int N, int T_local,
                         const float* d_thr, int32_t* d_ref_ok, int32_t* d_medoid, float* d_cos, float* d_emb,
                         void* stream) {
  if (c == nullptr || d_thr == nullptr || B < 0) return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (B == 0) return QC_OK;
  float* emb = d_emb;
  if (emb == nullptr) {
    DeviceGuard g(c->device);
    DevBuf& scratch = ws_for(c, stream ? static_cast<cudaStream_t>(stream) : c->stream)->emb;
    QC_CUDA(scratch.ensure(static_cast<size_t>(B) * N * c->dim * sizeof(float)));
    emb = static_cast<float*>(scratch.p);
  }
  int st = qc_embed_tsplit_dev(c, sp, d_tok, d_lens, B * N, T_local, emb, stream);
  if (st != QC_OK) return st;
  return qc_similarity_batch_dev(c, emb, B, N, c->dim, d_thr, d_ref_ok, d_medoid, d_cos, stream);
}

int qc_similarity_batch(qc_ctx* c, const float* emb, int B, int N, int D, const float* thr, int32_t* out_ref_ok,
                        int32_t* out_medoid, float* out_cos) {
  if (c == nullptr || emb == nullptr || thr == nullptr || B < 0 || D < 1) return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (B == 0) return QC_OK;
  DeviceGuard g(c->device);
  const size_t embb = static_cast<size_t>(B) * N * D * 4, thrb = static_cast<size_t>(B) * 4;
  const size_t cosb = static_cast<size_t>(B) * N * N * 4, ib = static_cast<size_t>(B) * 4;
  QC_CUDA(c->emb.ensure(embb));
  QC_CUDA(c->thr.ensure(thrb));
  QC_CUDA(c->refok.ensure(ib));
  QC_CUDA(c->medoid.ensure(ib));
  if (out_cos) QC_CUDA(c->cosm.ensure(cosb));
  QC_CUDA(c->h_in.ensure(pad256(embb) + pad256(thrb)));
  QC_CUDA(c->h_out.ensure(pad256(cosb) + 2 * pad256(ib)));
  size_t off = 0;
  int st;
  if ((st = h2d(c, c->emb.p, emb, embb, &off)) != QC_OK) return st;
  if ((st = h2d(c, c->thr.p, thr, thrb, &off)) != QC_OK) return st;
  st = qc_similarity_batch_dev(c, static_cast<float*>(c->emb.p), B, N, D, static_cast<float*>(c->thr.p),
                               static_cast<int32_t*>(c->refok.p), static_cast<int32_t*>(c->medoid.p),
                               out_cos ? static_cast<float*>(c->cosm.p) : nullptr, c->stream);
  if (st != QC_OK) return st;
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_ref_ok, c->refok.p, ib, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_medoid, c->medoid.p, ib, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_cos, c->cosm.p, cosb, &ooff, &pend)) != QC_OK) return st;
  QC_CUDA(cudaStreamSynchronize(c->stream));
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

int qc_decide_batch(qc_ctx* c, const int32_t* token_ids, const int32_t* lens, int B, int N, int T, const float* thr,
                    int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos, float* out_emb) {
  if (c == nullptr || token_ids == nullptr || lens == nullptr || thr == nullptr || B < 0 || T < 1)
    return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (T > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  if (B == 0) return QC_OK;
  const int R = B * N;
  for (int r = 0; r < R; ++r)
    if (lens[r] < 0 || lens[r] > T) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  const size_t tokb = static_cast<size_t>(R) * T * 4, lenb = static_cast<size_t>(R) * 4;
  const size_t thrb = static_cast<size_t>(B) * 4, ib = static_cast<size_t>(B) * 4;
  const size_t embb = static_cast<size_t>(R) * c->dim * 4, cosb = static_cast<size_t>(B) * N * N * 4;
  QC_CUDA(c->tok.ensure(tokb));
  QC_CUDA(c->lens.ensure(lenb));
  QC_CUDA(c->thr.ensure(thrb));
  QC_CUDA(c->emb.ensure(embb));
  QC_CUDA(c->refok.ensure(ib));
  QC_CUDA(c->medoid.ensure(ib));
  if (out_cos) QC_CUDA(c->cosm.ensure(cosb));
  QC_CUDA(c->h_in.ensure(pad256(tokb) + pad256(lenb) + pad256(thrb)));
  QC_CUDA(c->h_out.ensure(2 * pad256(ib) + pad256(cosb) + pad256(embb)));
  size_t off = 0;
  int st;
  if ((st = h2d(c, c->tok.p, token_ids, tokb, &off)) != QC_OK) return st;
  if ((st = h2d(c, c->lens.p, lens, lenb, &off)) != QC_OK) return st;
  if ((st = h2d(c, c->thr.p, thr, thrb, &off)) != QC_OK) return st;
  HintGuard hg(c);
  if ((st = stage_block_map(c, lens, R, T, &c->map, &c->h_map, &c->packed, c->stream)) != QC_OK) return st;
  st = qc_decide_batch_dev(c, static_cast<int32_t*>(c->tok.p), static_cast<int32_t*>(c->lens.p), B, N, T,
                           static_cast<float*>(c->thr.p), static_cast<int32_t*>(c->refok.p),
                           static_cast<int32_t*>(c->medoid.p), out_cos ? static_cast<float*>(c->cosm.p) : nullptr,
                           static_cast<float*>(c->emb.p), c->stream);
  if (st != QC_OK) return st;
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_ref_ok, c->refok.p, ib, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_medoid, c->medoid.p, ib, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_cos, c->cosm.p, cosb, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_emb, c->emb.p, embb, &ooff, &pend)) != QC_OK) return st;
  if ((st = check_token_flag(c)) != QC_OK) return st;
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

int qc_decide_batch_submit(qc_ctx* c, int slot, const int32_t* token_ids, const int32_t* lens, int B, int N, int T,
                           const float* thr, int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos) {
  if (c == nullptr || slot < 0 || slot >= QC_PIPELINE_SLOTS || token_ids == nullptr || lens == nullptr || thr == nullptr || B < 1 || T < 1)
    return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  if (T > QC_MAX_TOKENS) return QC_ERR_CONTEXT_LENGTH;
  Slot& sl = c->slots[slot];
  if (sl.busy) return QC_ERR_INVALID_INPUT;   // the previous batch in this slot has not been waited for
  const int R = B * N;
  for (int r = 0; r < R; ++r)
    if (lens[r] < 0 || lens[r] > T) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  if (sl.stream == nullptr) QC_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
  const size_t tokb = static_cast<size_t>(R) * T * 4, lenb = static_cast<size_t>(R) * 4;
  const size_t thrb = static_cast<size_t>(B) * 4, ib = static_cast<size_t>(B) * 4;
  const size_t embb = static_cast<size_t>(R) * c->dim * 4, cosb = static_cast<size_t>(B) * N * N * 4;
  QC_CUDA(sl.tok.ensure(tokb));
  QC_CUDA(sl.lens.ensure(lenb));
  QC_CUDA(sl.thr.ensure(thrb));
  QC_CUDA(sl.emb.ensure(embb));
  QC_CUDA(sl.refok.ensure(ib));
  QC_CUDA(sl.medoid.ensure(ib));
  if (out_cos) QC_CUDA(sl.cosm.ensure(cosb));
  QC_CUDA(sl.h_in.ensure(pad256(tokb) + pad256(lenb) + pad256(thrb)));
  QC_CUDA(sl.h_out.ensure(2 * pad256(ib) + pad256(cosb)));
  QC_CUDA(sl.h_flags.ensure(8 * sizeof(int32_t)));
  // stage / copy inputs on the slot's stream
  size_t off = 0;
  auto in = [&](void* dst, const void* src, size_t bytes) -> int {
    if (is_pinned(src)) {
      QC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sl.stream));
    } else {
      char* st = static_cast<char*>(sl.h_in.p) + off;
      memcpy(st, src, bytes);
      off += pad256(bytes);
      QC_CUDA(cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, sl.stream));
    }
    return QC_OK;
  };
  int st;
  if ((st = in(sl.tok.p, token_ids, tokb)) != QC_OK) return st;
  if ((st = in(sl.lens.p, lens, lenb)) != QC_OK) return st;
  if ((st = in(sl.thr.p, thr, thrb)) != QC_OK) return st;
  HintGuard hg(c);
  if ((st = stage_block_map(c, lens, R, T, &sl.map, &sl.h_map, &sl.packed, sl.stream)) != QC_OK) return st;
  st = encode_dev(c, static_cast<int32_t*>(sl.tok.p), static_cast<int32_t*>(sl.lens.p), R, T,
                  static_cast<float*>(sl.emb.p), sl.stream, &sl.part, nullptr, nullptr, err_block(c, slot));
  if (st != QC_OK) return st;
  st = similarity_dev(c, static_cast<float*>(sl.emb.p), B, N, c->dim, static_cast<float*>(sl.thr.p),
                      static_cast<int32_t*>(sl.refok.p), static_cast<int32_t*>(sl.medoid.p),
                      out_cos ? static_cast<float*>(sl.cosm.p) : nullptr, sl.stream, err_block(c, slot));
  if (st != QC_OK) return st;
  sl.pend.clear();
  size_t ooff = 0;
  auto outc = [&](void* user, const void* dsrc, size_t bytes) -> int {
    if (user == nullptr || bytes == 0) return QC_OK;
    if (is_pinned(user)) {
      QC_CUDA(cudaMemcpyAsync(user, dsrc, bytes, cudaMemcpyDeviceToHost, sl.stream));
    } else {
      char* stg = static_cast<char*>(sl.h_out.p) + ooff;
      ooff += pad256(bytes);
      QC_CUDA(cudaMemcpyAsync(stg, dsrc, bytes, cudaMemcpyDeviceToHost, sl.stream));
      sl.pend.push_back({user, stg, bytes});
    }
    return QC_OK;
  };
  if ((st = outc(out_ref_ok, sl.refok.p, ib)) != QC_OK) return st;
  if ((st = outc(out_medoid, sl.medoid.p, ib)) != QC_OK) return st;
  if ((st = outc(out_cos, sl.cosm.p, cosb)) != QC_OK) return st;
  QC_CUDA(cudaMemcpyAsync(sl.h_flags.p, err_block(c, slot), kErrWords * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
  sl.busy = true;
  return QC_OK;
}

int qc_decide_ragged_submit(qc_ctx* c, int slot, const int32_t* tokens, const int32_t* offsets, int B, int N,
                            const float* thr, int32_t* out_ref_ok, int32_t* out_medoid, float* out_cos) {
  if (c == nullptr || slot < 0 || slot >= QC_PIPELINE_SLOTS || offsets == nullptr || thr == nullptr || B < 1)
    return QC_ERR_INVALID_INPUT;
  if (N < 1 || N > QC_MAX_POOL) return QC_ERR_UNSUPPORTED;
  if (!c->has_encoder) return QC_ERR_NOT_CONFIGURED;
  Slot& sl = c->slots[slot];
  if (sl.busy) return QC_ERR_INVALID_INPUT;   // the previous batch in this slot has not been waited for
  const int R = B * N;
  if (offsets[R] < 0 || (offsets[R] > 0 && tokens == nullptr)) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  if (sl.stream == nullptr) QC_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
  const size_t thrb = static_cast<size_t>(B) * 4, ib = static_cast<size_t>(B) * 4;
  const size_t embb = static_cast<size_t>(R) * c->dim * 4, cosb = static_cast<size_t>(B) * N * N * 4;
  QC_CUDA(sl.thr.ensure(thrb));
  QC_CUDA(sl.emb.ensure(embb));
  QC_CUDA(sl.refok.ensure(ib));
  QC_CUDA(sl.medoid.ensure(ib));
  if (out_cos) QC_CUDA(sl.cosm.ensure(cosb));
  QC_CUDA(sl.h_out.ensure(2 * pad256(ib) + pad256(cosb) + pad256(thrb)));
  QC_CUDA(sl.h_flags.ensure(8 * sizeof(int32_t)));
  const RaggedBufs wb{&sl.tok, &sl.lens, &sl.map, &sl.packed, &sl.part, &sl.h_in, &sl.h_map};
  int st = ragged_common(c, wb, tokens, offsets, R, static_cast<float*>(sl.emb.p), sl.stream, err_block(c, slot));
  if (st != QC_OK) return st;
  {  // thresholds: staged at the tail of the slot's pinned output area unless the caller's buffer is pinned
    const void* src = thr;
    if (!is_pinned(thr)) {
      char* stg = static_cast<char*>(sl.h_out.p) + 2 * pad256(ib) + pad256(cosb);
      memcpy(stg, thr, thrb);
      src = stg;
    }
    QC_CUDA(cudaMemcpyAsync(sl.thr.p, src, thrb, cudaMemcpyHostToDevice, sl.stream));
  }
  st = similarity_dev(c, static_cast<float*>(sl.emb.p), B, N, c->dim, static_cast<float*>(sl.thr.p),
                      static_cast<int32_t*>(sl.refok.p), static_cast<int32_t*>(sl.medoid.p),
                      out_cos ? static_cast<float*>(sl.cosm.p) : nullptr, sl.stream, err_block(c, slot));
  if (st != QC_OK) return st;
  sl.pend.clear();
  size_t ooff = 0;
  auto outc = [&](void* user, const void* dsrc, size_t bytes) -> int {
    if (user == nullptr || bytes == 0) return QC_OK;
    if (is_pinned(user)) {
      QC_CUDA(cudaMemcpyAsync(user, dsrc, bytes, cudaMemcpyDeviceToHost, sl.stream));
    } else {
      char* stg = static_cast<char*>(sl.h_out.p) + ooff;
      ooff += pad256(bytes);
      QC_CUDA(cudaMemcpyAsync(stg, dsrc, bytes, cudaMemcpyDeviceToHost, sl.stream));
      sl.pend.push_back({user, stg, bytes});
    }
    return QC_OK;
  };
  if ((st = outc(out_ref_ok, sl.refok.p, ib)) != QC_OK) return st;
  if ((st = outc(out_medoid, sl.medoid.p, ib)) != QC_OK) return st;
  if ((st = outc(out_cos, sl.cosm.p, cosb)) != QC_OK) return st;
  QC_CUDA(cudaMemcpyAsync(sl.h_flags.p, err_block(c, slot), kErrWords * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
  sl.busy = true;
  return QC_OK;
}

int qc_decide_batch_wait(qc_ctx* c, int slot) {
  if (c == nullptr || slot < 0 || slot >= QC_PIPELINE_SLOTS) return QC_ERR_INVALID_INPUT;
  Slot& sl = c->slots[slot];
  if (!sl.busy) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(c->device);
  sl.busy = false;
  QC_CUDA(cudaStreamSynchronize(sl.stream));
  const int32_t* flags = static_cast<const int32_t*>(sl.h_flags.p);
  if (flags[0] != 0 || flags[4] != 0) {
    const bool hang = flags[0] != 0;
    if (hang) for (int i = 0; i < 4; ++i) c->hang[i] = flags[i];
    QC_CUDA(cudaMemsetAsync(err_block(c, slot), 0, kErrWords * sizeof(int32_t), sl.stream));
    QC_CUDA(cudaStreamSynchronize(sl.stream));
    return hang ? QC_ERR_KERNEL_TIMEOUT : QC_ERR_TOKEN_RANGE;
  }
  for (const PendingOut& p : sl.pend) memcpy(p.user, p.staged, p.bytes);
  sl.pend.clear();
  return QC_OK;
}

int qc_first_match_batch(qc_ctx* c, const float* query, const float* rows, int Q, int M, int D, const float* thr,
                         int32_t* out_index, float* out_cos) {
  if (c == nullptr || query == nullptr || thr == nullptr || out_index == nullptr || Q < 0 || M < 0 || D < 1)
    return QC_ERR_INVALID_INPUT;
  if (M > 0 && rows == nullptr) return QC_ERR_INVALID_INPUT;
  if (Q == 0) return QC_OK;
  DeviceGuard g(c->device);
  const size_t qb = static_cast<size_t>(Q) * D * 4, rb = static_cast<size_t>(Q) * M * D * 4;
  const size_t tb = static_cast<size_t>(Q) * 4, cb = static_cast<size_t>(Q) * M * 4;
  QC_CUDA(c->q.ensure(qb));
  QC_CUDA(c->rows.ensure(rb ? rb : 4));
  QC_CUDA(c->thr.ensure(tb));
  QC_CUDA(c->idx.ensure(tb));
  if (out_cos) QC_CUDA(c->cosm.ensure(cb ? cb : 4));
  QC_CUDA(c->h_in.ensure(pad256(qb) + pad256(rb) + pad256(tb)));
  QC_CUDA(c->h_out.ensure(pad256(tb) + pad256(cb)));
  size_t off = 0;
  int st;
  if ((st = h2d(c, c->q.p, query, qb, &off)) != QC_OK) return st;
  if ((st = h2d(c, c->rows.p, rows, rb, &off)) != QC_OK) return st;
  if ((st = h2d(c, c->thr.p, thr, tb, &off)) != QC_OK) return st;
  {
    KernelScope ks(c, QC_K_FIRST_MATCH, c->stream);
    QC_CUDA(qc::launch_first_match(static_cast<float*>(c->q.p), static_cast<float*>(c->rows.p), Q, M, D,
                                   static_cast<float*>(c->thr.p), static_cast<int32_t*>(c->idx.p),
                                   out_cos ? static_cast<float*>(c->cosm.p) : nullptr, c->stream));
  }
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_index, c->idx.p, tb, &ooff, &pend)) != QC_OK) return st;
  if ((st = d2h(c, out_cos, c->cosm.p, cb, &ooff, &pend)) != QC_OK) return st;
  QC_CUDA(cudaStreamSynchronize(c->stream));
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

// ---- CUDA graphs over the device entry points -------------------------------------------------------------------
// A caller whose batches are already in HBM (an inner loop of `_dev` calls on one stream) can record the loop once and
// replay it with one launch: no host work, and therefore no host jitter, between the kernels.  Everything the `_dev`
// entry points enqueue is capturable (kernel launches with their cluster / programmatic-dependency attributes); the
// workspaces must already have their size, i.e. the same calls must have run once before the capture.
struct qc_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int device = -1;
};

int qc_capture_begin(qc_ctx* c, void* stream) {
  if (c == nullptr || stream == nullptr) return QC_ERR_INVALID_INPUT;       // the legacy default stream cannot be captured
  DeviceGuard g(c->device);
  QC_CUDA(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeRelaxed));
  return QC_OK;
}

int qc_capture_end(qc_ctx* c, void* stream, qc_graph** out) {
  if (c == nullptr || stream == nullptr || out == nullptr) return QC_ERR_INVALID_INPUT;
  *out = nullptr;
  DeviceGuard g(c->device);
  cudaGraph_t graph = nullptr;
  QC_CUDA(cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &graph));
  qc_graph* q = new (std::nothrow) qc_graph();
  if (q == nullptr) {
    cudaGraphDestroy(graph);
    return QC_ERR_OUT_OF_MEMORY;
  }
  q->graph = graph;
  q->device = c->device;
  cudaError_t e = cudaGraphInstantiate(&q->exec, graph, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaGraphDestroy(graph);
    delete q;
    return cuda_status(e);
  }
  *out = q;
  return QC_OK;
}

int qc_graph_launch(qc_graph* q, void* stream) {
  if (q == nullptr || q->exec == nullptr) return QC_ERR_INVALID_INPUT;
  DeviceGuard g(q->device);
  QC_CUDA(cudaGraphLaunch(q->exec, static_cast<cudaStream_t>(stream)));
  return QC_OK;
}

void qc_graph_destroy(qc_graph* q) {
  if (q == nullptr) return;
  DeviceGuard g(q->device);
  if (q->exec) cudaGraphExecDestroy(q->exec);
  if (q->graph) cudaGraphDestroy(q->graph);
  delete q;
}

// ---- device-resident lesson rows -----------------------------------------------------
// LessonManager.find_similar_lesson/5 (reference lib/quoracle/agent/lesson_manager.ex:150-173) re-embeds every
// existing lesson (<= 100) for every new one; qc_first_match_batch already batches the comparison but still takes
// the [M x D] rows from the host on every call.  A lesson cache keeps an agent's rows in HBM: append / replace by
// index when a lesson is added or merged, reorder when the list is pruned, and a dedup call then uploads ONE query
// row (D floats) and reads back one index (+ M cosines when asked).
struct qc_lesson_cache {
  qc_ctx* ctx = nullptr;
  int capacity = 0, dim = 0, count = 0;
  float* rows = nullptr;      // [capacity x dim]
  float* spare = nullptr;     // gather target, swapped with rows
};

int qc_lesson_cache_create(qc_ctx* c, int capacity, int dim, qc_lesson_cache** out) {
  if (c == nullptr || out == nullptr || capacity < 1 || capacity > 65536 || dim < 1) return QC_ERR_INVALID_INPUT;
  *out = nullptr;
  DeviceGuard g(c->device);
  qc_lesson_cache* lc = new (std::nothrow) qc_lesson_cache();
  if (lc == nullptr) return QC_ERR_OUT_OF_MEMORY;
  lc->ctx = c;
  lc->capacity = capacity;
  lc->dim = dim;
  const size_t bytes = static_cast<size_t>(capacity) * dim * sizeof(float);
  cudaError_t e = cudaMalloc(&lc->rows, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&lc->spare, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (lc->rows) cudaFree(lc->rows);
    delete lc;
    return cuda_status(e);
  }
  *out = lc;
  return QC_OK;
}

void qc_lesson_cache_destroy(qc_lesson_cache* lc) {
  if (lc == nullptr) return;
  DeviceGuard g(lc->ctx->device);
  cudaStreamSynchronize(lc->ctx->stream);
  cudaFree(lc->rows);
  cudaFree(lc->spare);
  delete lc;
}

int qc_lesson_cache_count(const qc_lesson_cache* lc) { return lc ? lc->count : -1; }

int qc_lesson_cache_put(qc_lesson_cache* lc, int index, const float* rows, int nrows) {
  if (lc == nullptr || rows == nullptr || nrows < 1 || index < 0 || index > lc->count) return QC_ERR_INVALID_INPUT;
  if (index + nrows > lc->capacity) return QC_ERR_UNSUPPORTED;
  qc_ctx* c = lc->ctx;
  DeviceGuard g(c->device);
  const size_t bytes = static_cast<size_t>(nrows) * lc->dim * sizeof(float);
  QC_CUDA(c->h_in.ensure(pad256(bytes)));
  size_t off = 0;
  int st = h2d(c, lc->rows + static_cast<size_t>(index) * lc->dim, rows, bytes, &off);
  if (st != QC_OK) return st;
  QC_CUDA(cudaStreamSynchronize(c->stream));          // the staging area is reused by the next call
  if (index + nrows > lc->count) lc->count = index + nrows;
  return QC_OK;
}

int qc_lesson_cache_gather(qc_lesson_cache* lc, const int32_t* src_index, int n) {
  if (lc == nullptr || n < 0 || n > lc->capacity || (n > 0 && src_index == nullptr)) return QC_ERR_INVALID_INPUT;
  for (int i = 0; i < n; ++i)
    if (src_index[i] < 0 || src_index[i] >= lc->count) return QC_ERR_INVALID_INPUT;
  qc_ctx* c = lc->ctx;
  DeviceGuard g(c->device);
  const size_t rowb = static_cast<size_t>(lc->dim) * sizeof(float);
  for (int i = 0; i < n; ++i)                          // <= capacity device-to-device row copies on one stream
    QC_CUDA(cudaMemcpyAsync(lc->spare + static_cast<size_t>(i) * lc->dim, lc->rows + static_cast<size_t>(src_index[i]) * lc->dim,
                            rowb, cudaMemcpyDeviceToDevice, c->stream));
  float* t = lc->rows;
  lc->rows = lc->spare;
  lc->spare = t;
  lc->count = n;
  return QC_OK;
}

int qc_lesson_cache_match(qc_lesson_cache* lc, const float* query, float thr, int32_t* out_index, float* out_cos) {
  if (lc == nullptr || query == nullptr || out_index == nullptr) return QC_ERR_INVALID_INPUT;
  qc_ctx* c = lc->ctx;
  if (lc->count == 0) {
    *out_index = -1;
    return QC_OK;
  }
  DeviceGuard g(c->device);
  const size_t qb = static_cast<size_t>(lc->dim) * 4, cb = static_cast<size_t>(lc->count) * 4;
  QC_CUDA(c->q.ensure(qb + 256));
  QC_CUDA(c->idx.ensure(4));
  if (out_cos) QC_CUDA(c->cosm.ensure(cb));
  QC_CUDA(c->h_in.ensure(pad256(qb) + 256));
  QC_CUDA(c->h_out.ensure(pad256(cb) + 256));
  // the query row and its threshold ride in one staged copy: [D floats | thr]
  char* st = static_cast<char*>(c->h_in.p);
  memcpy(st, query, qb);
  memcpy(st + qb, &thr, 4);
  QC_CUDA(cudaMemcpyAsync(c->q.p, st, qb + 4, cudaMemcpyHostToDevice, c->stream));
  {
    KernelScope ks(c, QC_K_FIRST_MATCH, c->stream);
    QC_CUDA(qc::launch_first_match(static_cast<float*>(c->q.p), lc->rows, 1, lc->count, lc->dim,
                                   reinterpret_cast<float*>(static_cast<char*>(c->q.p) + qb), static_cast<int32_t*>(c->idx.p),
                                   out_cos ? static_cast<float*>(c->cosm.p) : nullptr, c->stream));
  }
  char* so = static_cast<char*>(c->h_out.p);
  QC_CUDA(cudaMemcpyAsync(so, c->idx.p, 4, cudaMemcpyDeviceToHost, c->stream));
  if (out_cos) QC_CUDA(cudaMemcpyAsync(so + 256, c->cosm.p, cb, cudaMemcpyDeviceToHost, c->stream));
  QC_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(out_index, so, 4);
  if (out_cos) memcpy(out_cos, so + 256, cb);
  return QC_OK;
}

int qc_chunk_mean(qc_ctx* c, const float* chunks, int G, int C, int D, float* out_mean) {
  if (c == nullptr || chunks == nullptr || out_mean == nullptr || G < 0 || C < 1 || D < 1) return QC_ERR_INVALID_INPUT;
  if (G == 0) return QC_OK;
  DeviceGuard g(c->device);
  const size_t ib = static_cast<size_t>(G) * C * D * 4, ob = static_cast<size_t>(G) * D * 4;
  QC_CUDA(c->rows.ensure(ib));
  QC_CUDA(c->q.ensure(ob));
  QC_CUDA(c->h_in.ensure(pad256(ib)));
  QC_CUDA(c->h_out.ensure(pad256(ob)));
  size_t off = 0;
  int st;
  if ((st = h2d(c, c->rows.p, chunks, ib, &off)) != QC_OK) return st;
  {
    KernelScope ks(c, QC_K_CHUNK_MEAN, c->stream);
    QC_CUDA(qc::launch_chunk_mean(static_cast<float*>(c->rows.p), G, C, D, static_cast<float*>(c->q.p), c->stream));
  }
  std::vector<PendingOut> pend;
  size_t ooff = 0;
  if ((st = d2h(c, out_mean, c->q.p, ob, &ooff, &pend)) != QC_OK) return st;
  QC_CUDA(cudaStreamSynchronize(c->stream));
  for (const PendingOut& p : pend) memcpy(p.user, p.staged, p.bytes);
  return QC_OK;
}

}  // extern "C"
