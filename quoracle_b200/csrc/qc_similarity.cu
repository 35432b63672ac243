// qc_similarity.cu -- K2 (finalise embeddings), K3 (N x N cosine + verdicts), lesson first-match,
// chunk mean.  All HBM-bandwidth-bound fp32 work: coalesced 16-byte loads, warp-shuffle reductions.
//
// K3 restates, per decision, the arithmetic of
//   ConsensusRules.cosine_similarity/2      reference lib/quoracle/actions/consensus_rules.ex:189-210
//   the row-0 threshold rule                reference lib/quoracle/actions/consensus_rules.ex:277-286
// and adds the north-star medoid vote (argmax_i mean_j cos(e_i, e_j), lowest index on ties),
// which has no reference counterpart.
#include "qc_common.cuh"
#include "ptx.cuh"

namespace qc {

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cos = dot / (|a| |b|), 0.0 when either norm is zero (reference consensus_rules.ex:205-209).
__device__ __forceinline__ float cosine_from(float dot, float na, float nb) {
  return (na == 0.f || nb == 0.f) ? 0.f : dot / (na * nb);
}

// ------------------------------------------------------------------------------------------
// K2: part[r, j, :] (j < ceil(len/TB)) -> mean over tokens -> L2 normalise -> emb[r, :]
// ------------------------------------------------------------------------------------------
template <int PER>
__global__ void __launch_bounds__(256) qc_finalize_kernel(const float* __restrict__ part,
                                                          const int32_t* __restrict__ lens, FinalizeDst dst,
                                                          const int32_t* __restrict__ row_start, int R, int bpc, int TB,
                                                          int D) {
  const int r = blockIdx.x;
  if (r >= R) return;
  ptx::griddep_wait();                 // K1's partial sums (programmatic dependent launch, see ptx.cuh)
  ptx::griddep_launch_dependents();
  const int len = lens[r];
  // dense: candidate r owns partial rows r*bpc .. (one per token block).  packed (row_start != nullptr): it owns
  // rows row_start[r] .. row_start[r+1]-1, one per 32-token chunk of the packed token stream.
  const int nb = row_start != nullptr ? row_start[r + 1] - row_start[r] : (len + TB - 1) / TB;
  const float* p = part + (row_start != nullptr ? static_cast<size_t>(row_start[r]) : static_cast<size_t>(r) * bpc) * D;
  // Block sums in fixed block order (j ascending per element), the PER elements of a thread side by side and four
  // blocks per trip so that a dozen independent loads are in flight (the kernel is latency-bound: 2048 small rows).
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) v[i] = 0.f;
#pragma unroll 4
  for (int j = 0; j < nb; ++j) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int d = threadIdx.x + i * 256;
      if (d < D) v[i] += p[static_cast<size_t>(j) * D + d];
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int d = threadIdx.x + i * 256;
    const float s = (d < D && len > 0) ? v[i] / static_cast<float>(len) : 0.f;
    v[i] = s;
    ss += s * s;
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float nrm = sqrtf(tot);
  const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
  // Local row r = (decision b, local candidate n).  In the gathered layout of a pool split across ranks it is
  // row (b, rank * n_local + n) of [B, n_total, D]; the same row goes to every destination (this rank's own
  // buffer and, over NVLink, each peer's): the all-gather is these stores.  Unsplit: n_local == n_total.
  const int b = r / dst.n_local;
  const size_t row = static_cast<size_t>(b) * dst.n_total + dst.row_offset + (r - b * dst.n_local);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int d = threadIdx.x + i * 256;
    if (d < D) {
      const float o = v[i] * inv;
      for (int q = 0; q < dst.count; ++q) dst.base[q][row * D + d] = o;
    }
  }
}

cudaError_t launch_finalize(const float* part, const int32_t* lens, const FinalizeDst& dst, const int32_t* row_start,
                            int R, int bpc, int TB, int dim, cudaStream_t stream, bool pdl) {
  if (R <= 0) return cudaSuccess;
  if (dst.count < 1 || dst.count > kMaxFinalizeDst || dst.n_local < 1) return cudaErrorInvalidValue;
  if (dim <= 512) return launch_kernel(qc_finalize_kernel<2>, dim3(R), dim3(256), 0, stream, pdl, part, lens, dst, row_start, R, bpc, TB, dim);
  if (dim <= 1024) return launch_kernel(qc_finalize_kernel<4>, dim3(R), dim3(256), 0, stream, pdl, part, lens, dst, row_start, R, bpc, TB, dim);
  if (dim <= 2048) return launch_kernel(qc_finalize_kernel<8>, dim3(R), dim3(256), 0, stream, pdl, part, lens, dst, row_start, R, bpc, TB, dim);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// Packed token stream for ragged batches: chunk c (`chunk` token slots) of the stream takes chunk_n[c] ids from
// token_ids + chunk_src[c]; the rest of the chunk is padding (-1 -> zero embedding row -> GELU(0) = 0).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qc_pack_tokens_kernel(const int32_t* __restrict__ token_ids,
                                                             const int32_t* __restrict__ chunk_src,
                                                             const int32_t* __restrict__ chunk_n,
                                                             int32_t* __restrict__ packed, int nchunks, int chunk) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int c = idx / chunk, i = idx - c * chunk;
  ptx::griddep_wait();                 // `packed` may still be read by the previous step's K1
  ptx::griddep_launch_dependents();
  if (c >= nchunks) return;
  packed[idx] = i < chunk_n[c] ? token_ids[static_cast<size_t>(chunk_src[c]) + i] : -1;
}

cudaError_t launch_pack_tokens(const int32_t* token_ids, const int32_t* chunk_src, const int32_t* chunk_n,
                               int32_t* packed, int nchunks, int chunk, cudaStream_t stream, bool pdl) {
  if (nchunks <= 0) return cudaSuccess;
  const long long slots = static_cast<long long>(nchunks) * chunk;
  if (slots > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int blocks = static_cast<int>((slots + 255) / 256);
  return launch_kernel(qc_pack_tokens_kernel, dim3(blocks), dim3(256), 0, stream, pdl, token_ids, chunk_src, chunk_n, packed,
                       nchunks, chunk);
}

// ------------------------------------------------------------------------------------------
// Candidates sharded on T across ranks (long-context path).
// partial: row r's block sums of THIS rank's token slice -> one fp32 partial sum, stored with the slice's token
//          count into slot `rank` of every rank's exchange buffer (peer memory; the all-gather is these stores).
// reduce:  emb[r] = L2normalise( (sum_p psum[p][r]) / (sum_p len[p][r]) ), p ascending -- the same value on
//          every rank, and the same quantity the unsplit encoder computes (fp32 association differs).
// ------------------------------------------------------------------------------------------
template <int PER>
__global__ void __launch_bounds__(256) qc_tsplit_partial_kernel(const float* __restrict__ part,
                                                                const int32_t* __restrict__ lens, TsplitDst dst, int R,
                                                                int bpc, int TB, int D) {
  const int r = blockIdx.x;
  if (r >= R) return;
  const int len = lens[r];
  const int nb = (len + TB - 1) / TB;
  const float* p = part + static_cast<size_t>(r) * bpc * D;
  const size_t row = (static_cast<size_t>(dst.rank) * dst.max_rows + r);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int d = threadIdx.x + i * 256;
    if (d < D) {
      float s = 0.f;
      for (int j = 0; j < nb; ++j) s += p[static_cast<size_t>(j) * D + d];  // fixed block order
      for (int q = 0; q < dst.count; ++q) dst.psum[q][row * D + d] = s;
    }
  }
  if (threadIdx.x == 0)
    for (int q = 0; q < dst.count; ++q) dst.lens[q][row] = len;
}

template <int PER>
__global__ void __launch_bounds__(256) qc_tsplit_reduce_kernel(const float* __restrict__ psum,
                                                               const int32_t* __restrict__ lens, float* __restrict__ emb,
                                                               int R, int world, int max_rows, int D) {
  const int r = blockIdx.x;
  if (r >= R) return;
  int total = 0;
  for (int p = 0; p < world; ++p) total += lens[static_cast<size_t>(p) * max_rows + r];
  float v[PER];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int d = threadIdx.x + i * 256;
    float s = 0.f;
    if (d < D) {
      for (int p = 0; p < world; ++p) s += psum[(static_cast<size_t>(p) * max_rows + r) * D + d];  // rank order
      s = total > 0 ? s / static_cast<float>(total) : 0.f;
    }
    v[i] = s;
    ss += s * s;
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float nrm = sqrtf(tot);
  const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int d = threadIdx.x + i * 256;
    if (d < D) emb[static_cast<size_t>(r) * D + d] = v[i] * inv;
  }
}

cudaError_t launch_tsplit_partial(const float* part, const int32_t* lens, const TsplitDst& dst, int R, int bpc, int TB,
                                  int dim, cudaStream_t stream) {
  if (R <= 0) return cudaSuccess;
  if (dst.count < 1 || dst.count > kMaxFinalizeDst) return cudaErrorInvalidValue;
  if (dim <= 512) qc_tsplit_partial_kernel<2><<<R, 256, 0, stream>>>(part, lens, dst, R, bpc, TB, dim);
  else if (dim <= 1024) qc_tsplit_partial_kernel<4><<<R, 256, 0, stream>>>(part, lens, dst, R, bpc, TB, dim);
  else if (dim <= 2048) qc_tsplit_partial_kernel<8><<<R, 256, 0, stream>>>(part, lens, dst, R, bpc, TB, dim);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_tsplit_reduce(const float* psum, const int32_t* lens, float* emb, int R, int world, int max_rows,
                                 int dim, cudaStream_t stream) {
  if (R <= 0) return cudaSuccess;
  if (dim <= 512) qc_tsplit_reduce_kernel<2><<<R, 256, 0, stream>>>(psum, lens, emb, R, world, max_rows, dim);
  else if (dim <= 1024) qc_tsplit_reduce_kernel<4><<<R, 256, 0, stream>>>(psum, lens, emb, R, world, max_rows, dim);
  else if (dim <= 2048) qc_tsplit_reduce_kernel<8><<<R, 256, 0, stream>>>(psum, lens, emb, R, world, max_rows, dim);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Cross-GPU handshake of the split-pool path (flags live in each rank's exported allocation).
// signal: after this rank's rows are stored everywhere, publish `epoch` in slot `rank` of every peer's flag
// array (system-scope release).  wait: spin until every peer's slot in the local array reached `epoch`
// (system-scope acquire); a bounded spin turns a missing peer into a watchdog record instead of a hang.
// ------------------------------------------------------------------------------------------
__global__ void qc_peer_signal_kernel(PeerFlags f, uint32_t epoch) {
  const int p = threadIdx.x;
  if (p >= f.world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.flags[p] + f.rank), "r"(epoch) : "memory");
}

__global__ void qc_peer_wait_kernel(PeerFlags f, uint32_t epoch, int32_t* dbg) {
  const int p = threadIdx.x;
  if (p >= f.world) return;
  const uint32_t* slot = f.flags[f.rank] + p;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
    if (static_cast<int32_t>(v - epoch) >= 0) break;
    if (clock64() - t0 > 6000000000LL) {          // ~3 s: a peer never arrived
      if (atomicCAS(dbg, 0, 1) == 0) {
        dbg[1] = 900 + p;
        dbg[2] = f.rank;
        dbg[3] = static_cast<int32_t>(epoch);
      }
      break;
    }
    __nanosleep(200);
  }
}

cudaError_t launch_peer_signal(const PeerFlags& f, uint32_t epoch, cudaStream_t stream) {
  qc_peer_signal_kernel<<<1, 32, 0, stream>>>(f, epoch);
  return cudaGetLastError();
}
cudaError_t launch_peer_wait(const PeerFlags& f, uint32_t epoch, int32_t* dbg, cudaStream_t stream) {
  qc_peer_wait_kernel<<<1, 32, 0, stream>>>(f, epoch, dbg);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Verdict tail shared by both K3 paths.  `c(i, j)` returns the cosine; runs on one thread.
// ------------------------------------------------------------------------------------------
template <class CosFn>
__device__ __forceinline__ void verdict_tail(int N, float thr, CosFn c, int32_t* ref_ok, int32_t* medoid) {
  int ok = 1;
  for (int i = 1; i < N; ++i) ok &= (c(0, i) >= thr) ? 1 : 0;
  int best = 0;
  float best_mean = 0.f;
  for (int i = 0; i < N; ++i) {
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += c(i, j);  // j ascending, diagonal included
    const float m = s / static_cast<float>(N);
    if (i == 0 || m > best_mean) {
      best_mean = m;
      best = i;
    }
  }
  if (ref_ok) *ref_ok = ok;
  if (medoid) *medoid = best;
}

// ------------------------------------------------------------------------------------------
// K3 path A: N <= 8.  WPD warps per decision (1 for large batches, 4 when the batch alone cannot fill
// the GPU); every lane streams 16-byte column chunks of all N rows and keeps the N(N+1)/2 partial dot
// products in registers.  The cross-lane reduction and the verdict tail are DISTRIBUTED through shared
// memory (lane p sums pair p in a fixed order, lane e forms cosine element e, lane i the row mean i):
// doing the tail redundantly on every lane cost as many instructions as the streaming loop itself.
// ------------------------------------------------------------------------------------------
template <int N, bool VEC, int WPD>
__global__ void __launch_bounds__(256) qc_sim_small_kernel(const float* __restrict__ emb, int B, int D,
                                                           const float* __restrict__ thr,
                                                           int32_t* __restrict__ ref_ok,
                                                           int32_t* __restrict__ medoid, float* __restrict__ cos) {
  constexpr int P = N * (N + 1) / 2;
  constexpr int DPB = 8 / WPD;             // decisions per 8-warp block
  constexpr int GT = 32 * WPD;             // threads cooperating on one decision
  __shared__ float red[DPB][P][GT + 1];
  __shared__ float tot[DPB][P];
  __shared__ float seg_sum[DPB][P][WPD];
  __shared__ float cs[DPB][N * N];
  __shared__ float rmean[DPB][N];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int ds = warp / WPD;               // decision slot in this block
  const int slice = warp % WPD;            // which part of the columns this warp streams
  const int tg = slice * 32 + lane;        // thread index within the decision group
  const int b = blockIdx.x * DPB + ds;
  const bool live = b < B;
  ptx::griddep_wait();                 // K2's rows (programmatic dependent launch, see ptx.cuh)
  ptx::griddep_launch_dependents();

  float acc[P];
#pragma unroll
  for (int p = 0; p < P; ++p) acc[p] = 0.f;
  if (live) {
    const float* x = emb + static_cast<size_t>(b) * N * D;
    if constexpr (VEC) {
      for (int col = tg * 4; col < D; col += GT * 4) {
        float4 v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = ldg_stream4(x + static_cast<size_t>(i) * D + col);
        int p = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
          for (int j = i; j < N; ++j, ++p) {
            acc[p] = fmaf(v[i].x, v[j].x, acc[p]);
            acc[p] = fmaf(v[i].y, v[j].y, acc[p]);
            acc[p] = fmaf(v[i].z, v[j].z, acc[p]);
            acc[p] = fmaf(v[i].w, v[j].w, acc[p]);
          }
        }
      }
    } else {
      for (int col = tg; col < D; col += GT) {
        float v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __ldg(x + static_cast<size_t>(i) * D + col);
        int p = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
          for (int j = i; j < N; ++j, ++p) acc[p] = fmaf(v[i], v[j], acc[p]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) red[ds][p][tg] = acc[p];
  __syncthreads();
  // pair p: fixed-order sum of the GT per-thread partials.  One warp per decision: lane p walks its 32 partials.  Four
  // warps per decision (small batches, latency matters): the 128 partials of a pair are cut into NSEG segments summed by
  // different threads, then the segment sums are added in order -- a fixed tree, so results stay run-to-run identical,
  // with a dependent chain of 32 + 4 adds instead of 128.
  if constexpr (WPD == 1) {
    for (int p = tg; p < P; p += GT) {
      float s = 0.f;
#pragma unroll 8
      for (int t = 0; t < GT; ++t) s += red[ds][p][t];
      tot[ds][p] = s;
    }
  } else {
    constexpr int NSEG = GT / 32;
    for (int u = tg; u < P * NSEG; u += GT) {
      const int p = u / NSEG, seg = u - p * NSEG;
      float s = 0.f;
#pragma unroll 8
      for (int t = 0; t < 32; ++t) s += red[ds][p][seg * 32 + t];
      seg_sum[ds][p][seg] = s;
    }
    __syncthreads();
    for (int p = tg; p < P; p += GT) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NSEG; ++q) s += seg_sum[ds][p][q];
      tot[ds][p] = s;
    }
  }
  __syncthreads();
  // cosine element e = (i, j); pair index of (i <= j) is i*N - i(i-1)/2 + (j - i)
  for (int e = tg; e < N * N; e += GT) {
    const int i = e / N, j = e - i * N;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const float dot = tot[ds][lo * N - lo * (lo - 1) / 2 + (hi - lo)];
    const float ni = sqrtf(tot[ds][i * N - i * (i - 1) / 2]);
    const float nj = sqrtf(tot[ds][j * N - j * (j - 1) / 2]);
    const float c = cosine_from(dot, ni, nj);
    cs[ds][e] = c;
    if (live && cos != nullptr) cos[static_cast<size_t>(b) * N * N + e] = c;
  }
  __syncthreads();
  if (tg < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) s += cs[ds][tg * N + j];   // j ascending, diagonal included
    rmean[ds][tg] = s / static_cast<float>(N);
  }
  __syncthreads();
  if (tg == 0 && live) {
    const float th = thr[b];
    int ok = 1;
#pragma unroll
    for (int i = 1; i < N; ++i) ok &= (cs[ds][i] >= th) ? 1 : 0;
    int best = 0;
    float best_mean = rmean[ds][0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
      if (rmean[ds][i] > best_mean) {
        best_mean = rmean[ds][i];
        best = i;
      }
    }
    if (ref_ok) ref_ok[b] = ok;
    if (medoid) medoid[b] = best;
  }
}

// ------------------------------------------------------------------------------------------
// K3 path B: 8 < N <= 64.  The Gram matrix X X^T is GEMM-shaped and, at 2 N^2 D flop per 4 N D bytes,
// far past what the fp32 FMA pipe sustains against HBM (N = 64: 32 flop/B), so it runs on the tensor
// cores: warp-level m16n8k8 TF32 MMAs with the 3xTF32 split (x = hi + lo; hi.hi + hi.lo + lo.hi), which
// keeps the dot products at fp32 accuracy (error ~1e-7, the verdict tolerance is 1e-3).  One CTA per
// decision; 64-column chunks of all rows are staged in shared memory with cp.async (double buffered,
// padded rows: fragment loads are bank-conflict free); a warp owns one 16-row block x up to four
// 8-column tiles, small pools split the k range over warps instead.
// ------------------------------------------------------------------------------------------
constexpr int kSimDC = 64;              // columns per chunk
constexpr int kSimRS = kSimDC + 4;      // padded row stride (floats)
constexpr int kSimThreads = 256;

template <bool VEC>
__device__ __forceinline__ void sim_load_chunk(float* buf, const float* x, int N, int Np, int D, int c0) {
  constexpr int G = kSimDC / 4;
  for (int idx = threadIdx.x; idx < Np * G; idx += kSimThreads) {
    const int row = idx / G;
    const int g = idx - row * G;
    const int col = c0 + 4 * g;
    float* dst = buf + row * kSimRS + 4 * g;
    if constexpr (VEC) {
      const bool in = (row < N) && (col < D);
      ptx::cp_async_16(ptx::smem_u32(dst), in ? (x + static_cast<size_t>(row) * D + col) : x, in ? 16u : 0u);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool in = (row < N) && (col + e < D);
        dst[e] = in ? __ldg(x + static_cast<size_t>(row) * D + col + e) : 0.f;
      }
    }
  }
  if constexpr (VEC) ptx::cp_async_commit();
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rest = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool VEC>
__global__ void __launch_bounds__(kSimThreads) qc_sim_large_kernel(const float* __restrict__ emb, int B, int N, int D,
                                                                   const float* __restrict__ thr,
                                                                   int32_t* __restrict__ ref_ok,
                                                                   int32_t* __restrict__ medoid,
                                                                   float* __restrict__ cos) {
  extern __shared__ __align__(16) float sm[];
  const int Np = (N + 15) & ~15;                 // 16, 32, 48, 64
  float* buf0 = sm;
  float* buf1 = sm + 64 * kSimRS;
  float* gram = sm + 2 * 64 * kSimRS;            // [64][65]; also k-slice partials for Np == 16

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, tq = lane & 3;        // mma fragment coordinates
  // warp -> (row block rb, first column tile cb0, tiles ncb, k-slice ks of nks)
  int rb, cb0, ncb, ks = 0, nks = 1;
  if (Np == 16) { rb = 0; cb0 = warp & 1; ncb = 1; ks = warp >> 1; nks = 4; }
  else if (Np == 32) { rb = warp >> 2; cb0 = warp & 3; ncb = 1; }
  else if (Np == 48) { rb = warp >> 1; cb0 = (warp & 1) * 3; ncb = warp < 6 ? 3 : 0; }
  else { rb = warp >> 1; cb0 = (warp & 1) * 4; ncb = 4; }

  ptx::griddep_wait();                 // K2's rows (programmatic dependent launch, see ptx.cuh)
  ptx::griddep_launch_dependents();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* x = emb + static_cast<size_t>(b) * N * D;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;

    const int nchunks = (D + kSimDC - 1) / kSimDC;
    sim_load_chunk<VEC>(buf0, x, N, Np, D, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
      const float* cur = (ch & 1) ? buf1 : buf0;
      float* nxt = (ch & 1) ? buf0 : buf1;
      if (ch + 1 < nchunks) {
        sim_load_chunk<VEC>(nxt, x, N, Np, D, (ch + 1) * kSimDC);
        if constexpr (VEC) ptx::cp_async_wait<1>();
      } else {
        if constexpr (VEC) ptx::cp_async_wait<0>();
      }
      __syncthreads();
      if (ncb > 0) {
        const float* arow0 = cur + (rb * 16 + g) * kSimRS;
        const float* arow1 = arow0 + 8 * kSimRS;
        for (int k0 = ks * 8; k0 < kSimDC; k0 += 8 * nks) {
          uint32_t ah[4], al[4];
          split_tf32(arow0[k0 + tq], ah[0], al[0]);
          split_tf32(arow1[k0 + tq], ah[1], al[1]);
          split_tf32(arow0[k0 + tq + 4], ah[2], al[2]);
          split_tf32(arow1[k0 + tq + 4], ah[3], al[3]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < ncb) {
              const float* brow = cur + ((cb0 + c) * 8 + g) * kSimRS;
              uint32_t bh[2], bl[2];
              split_tf32(brow[k0 + tq], bh[0], bl[0]);
              split_tf32(brow[k0 + tq + 4], bh[1], bl[1]);
              mma_tf32(acc[c], al, bh);      // small terms first
              mma_tf32(acc[c], ah, bl);
              mma_tf32(acc[c], ah, bh);
            }
          }
        }
      }
      __syncthreads();
    }
    // accumulators -> Gram matrix in shared memory.  c0:(g, 2tq) c1:(g, 2tq+1) c2:(g+8, 2tq) c3:(g+8, 2tq+1)
    if (nks > 1) {
      float* red = buf0;                       // [nks][16][17]
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = g + (e >> 1) * 8, cc = cb0 * 8 + 2 * tq + (e & 1);
        red[(ks * 16 + r) * 17 + cc] = acc[0][e];
      }
      __syncthreads();
      if (t < 256) {
        const int r = t >> 4, cc = t & 15;
        float s = 0.f;
        for (int k = 0; k < nks; ++k) s += red[(k * 16 + r) * 17 + cc];   // fixed order
        gram[r * 65 + cc] = s;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < ncb) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = rb * 16 + g + (e >> 1) * 8, cc = (cb0 + c) * 8 + 2 * tq + (e & 1);
            gram[r * 65 + cc] = acc[c][e];
          }
        }
      }
    }
    __syncthreads();
    // cosine in place (needs the diagonal first): stash norms, then overwrite
    float* nrm = buf1;  // Np floats
    if (t < Np) nrm[t] = sqrtf(gram[t * 65 + t]);
    __syncthreads();
    for (int idx = t; idx < N * N; idx += kSimThreads) {
      const int i = idx / N, j = idx - i * N;
      // G is symmetric up to rounding: use the upper-triangular entry for both (i, j) and (j, i)
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      const float c = cosine_from(gram[lo * 65 + hi], nrm[i], nrm[j]);
      if (cos != nullptr) cos[static_cast<size_t>(b) * N * N + idx] = c;
      buf0[idx] = c;                            // N*N <= 4096 floats fit in buf0 (64 * 68)
    }
    __syncthreads();
    if (t < N) {
      float s = 0.f;
      for (int j = 0; j < N; ++j) s += buf0[t * N + j];   // j ascending, diagonal included
      nrm[64 + t] = s / static_cast<float>(N);
    }
    __syncthreads();
    if (t == 0) {
      const float th = thr[b];
      int ok = 1;
      for (int i = 1; i < N; ++i) ok &= (buf0[i] >= th) ? 1 : 0;
      int best = 0;
      float best_mean = nrm[64];
      for (int i = 1; i < N; ++i)
        if (nrm[64 + i] > best_mean) {
          best_mean = nrm[64 + i];
          best = i;
        }
      if (ref_ok) ref_ok[b] = ok;
      if (medoid) medoid[b] = best;
    }
    __syncthreads();
  }
}

constexpr int kSimLargeSmem = (2 * 64 * kSimRS + 64 * 65) * static_cast<int>(sizeof(float));

template <int N>
static cudaError_t launch_small(const float* emb, int B, int D, const float* thr, int32_t* ref_ok, int32_t* medoid,
                                float* cos, bool vec, int num_sms, cudaStream_t stream, bool pdl) {
  // one warp per decision needs >= ~16 warps per SM in flight to cover HBM latency; below that, split the
  // columns of each decision over four warps
  const bool wide = B < num_sms * 16;
  if (wide) {
    const int grid = (B + 1) / 2;
    if (vec) return launch_kernel(qc_sim_small_kernel<N, true, 4>, dim3(grid), dim3(256), 0, stream, pdl, emb, B, D, thr, ref_ok, medoid, cos);
    return launch_kernel(qc_sim_small_kernel<N, false, 4>, dim3(grid), dim3(256), 0, stream, pdl, emb, B, D, thr, ref_ok, medoid, cos);
  } else {
    const int grid = (B + 7) / 8;
    if (vec) return launch_kernel(qc_sim_small_kernel<N, true, 1>, dim3(grid), dim3(256), 0, stream, pdl, emb, B, D, thr, ref_ok, medoid, cos);
    return launch_kernel(qc_sim_small_kernel<N, false, 1>, dim3(grid), dim3(256), 0, stream, pdl, emb, B, D, thr, ref_ok, medoid, cos);
  }
}

// Shared-memory opt-in (51 KB) of the mma.sync kernel on the current device (called by qc_ctx_create).
cudaError_t similarity_prepare_device() {
  cudaError_t e = cudaFuncSetAttribute(qc_sim_large_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSimLargeSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(qc_sim_large_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSimLargeSmem);
}

cudaError_t launch_similarity(const float* emb, int B, int N, int D, const float* thr, int32_t* ref_ok,
                              int32_t* medoid, float* cos, int num_sms, cudaStream_t stream, bool pdl) {
  if (B <= 0) return cudaSuccess;
  if (N < 1 || N > 64 || D < 1) return cudaErrorInvalidValue;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(emb) & 15) == 0);
  switch (N) {
    case 1: return launch_small<1>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 2: return launch_small<2>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 3: return launch_small<3>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 4: return launch_small<4>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 5: return launch_small<5>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 6: return launch_small<6>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 7: return launch_small<7>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    case 8: return launch_small<8>(emb, B, D, thr, ref_ok, medoid, cos, vec, num_sms, stream, pdl);
    default: break;
  }
  int grid = B < num_sms * 4 ? B : num_sms * 4;  // four resident CTAs per SM (51 KB each)
  if (vec) return launch_kernel(qc_sim_large_kernel<true>, dim3(grid), dim3(kSimThreads), kSimLargeSmem, stream, pdl, emb, B, N, D, thr, ref_ok, medoid, cos);
  return launch_kernel(qc_sim_large_kernel<false>, dim3(grid), dim3(kSimThreads), kSimLargeSmem, stream, pdl, emb, B, N, D, thr, ref_ok, medoid, cos);
}

// ------------------------------------------------------------------------------------------
// Lesson dedup (reference lib/quoracle/agent/lesson_manager.ex:150-173): first row m with
// cos(query, row_m) >= thr.  One CTA per query, one warp per row (round robin).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qc_first_match_kernel(const float* __restrict__ query,
                                                             const float* __restrict__ rows, int Q, int M, int D,
                                                             const float* __restrict__ thr,
                                                             int32_t* __restrict__ out_index,
                                                             float* __restrict__ out_cos) {
  __shared__ int first;
  const int q = blockIdx.x;
  if (q >= Q) return;
  if (threadIdx.x == 0) first = 0x7fffffff;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qv = query + static_cast<size_t>(q) * D;
  const float th = thr[q];
  for (int m = warp; m < M; m += 8) {
    const float* rv = rows + (static_cast<size_t>(q) * M + m) * D;
    float dot = 0.f, nq = 0.f, nr = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float a = __ldg(qv + d), b = __ldg(rv + d);
      dot = fmaf(a, b, dot);
      nq = fmaf(a, a, nq);
      nr = fmaf(b, b, nr);
    }
    dot = warp_sum(dot);
    nq = warp_sum(nq);
    nr = warp_sum(nr);
    const float c = cosine_from(dot, sqrtf(nq), sqrtf(nr));
    if (lane == 0) {
      if (out_cos) out_cos[static_cast<size_t>(q) * M + m] = c;
      if (c >= th) atomicMin(&first, m);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_index[q] = first == 0x7fffffff ? -1 : first;
}

cudaError_t launch_first_match(const float* query, const float* rows, int Q, int M, int D, const float* thr,
                               int32_t* out_index, float* out_cos, cudaStream_t stream) {
  if (Q <= 0) return cudaSuccess;
  qc_first_match_kernel<<<Q, 256, 0, stream>>>(query, rows, Q, M, D, thr, out_index, out_cos);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Chunk mean (reference lib/quoracle/models/embeddings.ex:371-383): element-wise unweighted mean,
// chunks added in order, NOT re-normalised.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qc_chunk_mean_kernel(const float* __restrict__ chunks, int G, int C, int D,
                                                            float* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(G) * D) return;
  const size_t g = idx / D, d = idx - g * D;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += chunks[(g * C + c) * D + d];
  out[idx] = s / static_cast<float>(C);
}

cudaError_t launch_chunk_mean(const float* chunks, int G, int C, int D, float* out, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(G) * D;
  if (total == 0) return cudaSuccess;
  qc_chunk_mean_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(chunks, G, C, D, out);
  return cudaGetLastError();
}

}  // namespace qc
