// qc_similarity_tc.cu -- K3 for pools of 9..64 candidates on the 5th-gen tensor cores  (sm_100a)
//
// The Gram matrix X X^T of a decision is GEMM-shaped and, at 2 N^2 D flop per 4 N D bytes, beyond what
// the fp32 FMA pipe (or the legacy mma.sync path, profiles/sim_sweep_r1.md) sustains against HBM.  Here:
//   * a tile is 128 consecutive rows of the [B*N x D] embedding matrix = floor(128/N) whole decisions;
//     TMA streams it in 32-column fp32 chunks (128 rows x 128 B, 128-byte swizzle) through a 4-deep ring;
//   * four splitter warps turn each raw chunk into a TF32 "hi" tile (low mantissa bits cleared) and a
//     "lo" tile (x - hi), written with the same swizzled offsets, and publish them to the async proxy;
//   * one elected lane issues tcgen05.mma kind::tf32, M = 128, N = 256 (B = [hi; lo]), K = 8:
//     [HH | HL] += H [H; L]^T into one 256-column TMEM accumulator (3xTF32 with the symmetric term taken from the transpose:
//     G = HH + HL + HL^T, fp32-accurate to ~1e-7 -- the verdict tolerance is 1e-3);
//   * four epilogue warps read the accumulator rows back with tcgen05.ld, keep only the columns of the
//     row's own decision (the off-diagonal blocks of the 128 x 128 tile pair different decisions and are
//     discarded), and finish cosine / row mean / reference verdict / medoid vote through shared memory.
// Arithmetic per 16 KB chunk: 4 MMAs of 128x256x8; the kernel is co-limited by HBM and by shared-memory
// bandwidth (each MMA reads 8 KB of operands).
//
// Cosine and verdict definitions are those of qc_similarity.cu (reference
// lib/quoracle/actions/consensus_rules.ex:189-210,277-286; medoid: SURVEY.md section 8(a) row A13).
#include "qc_common.cuh"
#include "ptx.cuh"

namespace qc {

using namespace ptx;

namespace {

constexpr int kTcThreads = 448;             // 4 epilogue + 8 splitter + 1 TMA + 1 MMA warps
constexpr int kSplitWarps = 8;
constexpr int kRawStages = 4;
constexpr int kHlStages = 2;
constexpr int kChunkBytes = 128 * 128;       // 128 rows x 32 fp32
constexpr int kWin = 65;                     // row stride of the per-decision windows (N <= 64)
constexpr int kTcSmem = 1024 + kRawStages * kChunkBytes + kHlStages * 2 * kChunkBytes + 2 * 128 * kWin * 4 + 128 * 4 * 2 + 256 + 2560;   // tail: barriers + over-read padding

// kind::tf32, fp32 accumulate, both operands K-major: c_format=1 (bit 4), a/b_format=2 (bits 7, 10), N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Round-to-nearest (ties away) to the 10-bit TF32 mantissa; finite inputs only (the embeddings are).
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 1)
qc_sim_tc_kernel(const __grid_constant__ CUtensorMap x_tmap,   // emb as [B*N rows x D cols] fp32, box 32 x 128
                 int B, int N, int D, const float* __restrict__ thr, int32_t* __restrict__ ref_ok,
                 int32_t* __restrict__ medoid, float* __restrict__ cos, int32_t* __restrict__ dbg,
                 unsigned long long* __restrict__ stall) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* raw = smem;                                        // kRawStages x 16 KB
  uint8_t* hl = raw + kRawStages * kChunkBytes;               // kHlStages x (hi 16 KB, lo 16 KB)
  float* whh = reinterpret_cast<float*>(hl + kHlStages * 2 * kChunkBytes);   // [128][kWin] HH window
  float* whl = whh + 128 * kWin;                                             // [128][kWin] HL window
  float* diag = whl + 128 * kWin;                                            // [128] Gram diagonal
  float* rmean = diag + 128;                                                 // [128] row means
  uint64_t* bars = reinterpret_cast<uint64_t*>(rmean + 128);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + kRawStages;
  uint64_t* hl_full = raw_empty + kRawStages;
  uint64_t* hl_empty = hl_full + kHlStages;
  uint64_t* acc_full = hl_empty + kHlStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q = 128 / N;                                      // whole decisions per tile
  const int ntiles = (B + q - 1) / q;
  const int nchunks = (D + 31) / 32;

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&x_tmap);
    for (int s = 0; s < kRawStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], kSplitWarps);      // one arrival per splitter warp
    }
    for (int u = 0; u < kHlStages; ++u) {
      mbar_init(&hl_full[u], kSplitWarps);
      mbar_init(&hl_empty[u], 1);       // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);      // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 13) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== epilogue =========================================================
    const int i = warp * 32 + lane;                           // tile row == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const int d = i / N;                                      // decision slot of this row
    const int r = i - d * N;                                  // row inside the decision
    const int base = d * N;                                   // first tile column of the decision
    const bool row_live = d < q;
    uint32_t acc_iter = 0;
    const long long t_role = clock64();
    long long w_acc = 0, t_copy = 0, t_loop = 0, t_vote = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++acc_iter) {
      const uint32_t as = acc_iter & 1u;
      const int b = tile * q + d;
      const bool live = row_live && b < B;
      const float th = live ? thr[b] : 0.f;                   // in flight while the tile's MMAs finish
      w_acc += mbar_wait(&acc_full[as], (acc_iter >> 1) & 1u, dbg, 700);
      long long tp = clock64();
      tcgen05_fence_after();
      // accumulator rows -> shared windows (only the columns of the row's own decision; 32-column groups
      // that no row of this warp needs are skipped warp-uniformly)
      const int col_lo = ((warp * 32) / N) * N;
      const int col_hi = ((warp * 32 + 31) / N) * N + N;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float* dstw = (half == 0 ? whh : whl) + i * kWin - base;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          if (c0 + 32 <= col_lo || c0 >= col_hi) continue;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + as * 256 + half * 128 + c0, v);
          tmem_ld_wait();
          if (row_live) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int c = c0 + k - base;
              if (c >= 0 && c < N) dstw[c0 + k] = __uint_as_float(v[k]);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      named_bar_sync(1, 128);
      t_copy += clock64() - tp; tp = clock64();
      // Gram diagonal G_ii = HH_ii + 2 HL_ii -> row norm
      if (row_live) diag[i] = sqrtf((whl[i * kWin + r] + whl[i * kWin + r]) + whh[i * kWin + r]);
      named_bar_sync(1, 128);
      int ok = 1;
      if (live) {
        const float ni = diag[i];
        float s = 0.f;
        // G_rj = HH_rj + (HL_rj + HL_jr).  HH = H H^T comes out of the tensor core bitwise symmetric (exact
        // tf32 products, same k order for (r, j) and (j, r)), and the bracket is symmetric by commutativity, so
        // cos(r, j) == cos(j, r) exactly (asserted in tests/test_gpu_similarity.py).
        const float* hh_row = whh + i * kWin;
        const float* hl_row = whl + i * kWin;
        const float* hl_col = whl + base * kWin + r;
        const float* nrm = diag + base;
        float* cos_row = cos != nullptr ? cos + (static_cast<size_t>(b) * N + r) * N : nullptr;
        // Blocks of 8 columns, all loads of a block issued before use (shared memory is kept busy by the MMA
        // operand reads, so its latency is long here).  Loads past column N - 1 stay inside the padded
        // shared-memory carve-up and their results are discarded.
        for (int j0 = 0; j0 < N; j0 += 8) {
          float up[8], lr[8], lc[8], nj[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            up[k] = hh_row[j0 + k];
            lr[k] = hl_row[j0 + k];
            lc[k] = hl_col[(j0 + k) * kWin];
            nj[k] = nrm[j0 + k];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = j0 + k;
            if (j < N) {
              const float g = (lr[k] + lc[k]) + up[k];
              const float nn = ni * nj[k];
              const float c = nn == 0.f ? 0.f : __fdividef(g, nn);
              if (j > 0 && !(c >= th)) ok = 0;                // only row 0's verdict is used
              s += c;                                         // j ascending, diagonal included
              if (cos_row != nullptr) cos_row[j] = c;
            }
          }
        }
        rmean[i] = s / static_cast<float>(N);
      }
      named_bar_sync(1, 128);
      t_loop += clock64() - tp; tp = clock64();
      if (live && r == 0) {
        int best = 0;
        float best_mean = rmean[i];
        for (int j0 = 1; j0 < N; j0 += 8) {
          float m[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = rmean[i + j0 + k];   // over-reads land in the padding
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (j0 + k < N && m[k] > best_mean) {             // strict: ties keep the lowest index
              best_mean = m[k];
              best = j0 + k;
            }
        }
        if (ref_ok) ref_ok[b] = ok;
        if (medoid) medoid[b] = best;
      }
      named_bar_sync(1, 128);                                 // windows are reused by the next tile
      t_vote += clock64() - tp; 
    }
    if (stall != nullptr && threadIdx.x == 0) {
      stall[blockIdx.x * 16 + 4] = clock64() - t_role;
      stall[blockIdx.x * 16 + 5] = w_acc;
      stall[blockIdx.x * 16 + 11] = t_copy;
      stall[blockIdx.x * 16 + 12] = t_loop;
      stall[blockIdx.x * 16 + 13] = t_vote;
    }
  } else if (warp < 4 + kSplitWarps) {
    // ===================== splitters: raw fp32 chunk -> TF32 hi / lo tiles ====================
    const int p = threadIdx.x - 128;                          // 0..255
    uint32_t ci = 0;
    const long long t_role = clock64();
    long long w_raw = 0, w_hl = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t s = ci % kRawStages, u = ci % kHlStages;
        w_raw += mbar_wait(&raw_full[s], (ci / kRawStages) & 1u, dbg, 710);
        w_hl += mbar_wait(&hl_empty[u], ((ci / kHlStages) & 1u) ^ 1u, dbg, 711);
        const float4* src = reinterpret_cast<const float4*>(raw + s * kChunkBytes);
        float4* dh = reinterpret_cast<float4*>(hl + u * 2 * kChunkBytes);
        float4* dl = reinterpret_cast<float4*>(hl + u * 2 * kChunkBytes + kChunkBytes);
        // 1024 16-byte pieces, consecutive per warp; all loads first so their latencies overlap
        float4 x[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) x[it] = src[it * (32 * kSplitWarps) + p];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int idx = it * (32 * kSplitWarps) + p;
          float4 h, l;
          h.x = tf32_hi(x[it].x); h.y = tf32_hi(x[it].y); h.z = tf32_hi(x[it].z); h.w = tf32_hi(x[it].w);
          // the residual is exact in fp32; the tensor core drops its low bits itself (2^-21 relative overall)
          l.x = x[it].x - h.x; l.y = x[it].y - h.y; l.z = x[it].z - h.z; l.w = x[it].w - h.w;
          dh[idx] = h;
          dl[idx] = l;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&hl_full[u]);
          mbar_arrive(&raw_empty[s]);
        }
      }
    }
    if (stall != nullptr && p == 0) {
      stall[blockIdx.x * 16 + 6] = clock64() - t_role;
      stall[blockIdx.x * 16 + 7] = w_raw;
      stall[blockIdx.x * 16 + 10] = w_hl;
    }
  } else if (warp == 12) {
    // ===================== TMA producer (warp-convergent, elected issue) =====================
    uint32_t ci = 0;
    const long long t_role = clock64();
    long long w_free = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int row0 = tile * q * N;
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t s = ci % kRawStages;
        w_free += mbar_wait(&raw_empty[s], ((ci / kRawStages) & 1u) ^ 1u, dbg, 720);
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], kChunkBytes);
          tma_load_2d(raw + s * kChunkBytes, &x_tmap, &raw_full[s], ch * 32, row0);
        }
        __syncwarp();
      }
    }
    if (stall != nullptr && lane == 0) {
      stall[blockIdx.x * 16 + 8] = clock64() - t_role;
      stall[blockIdx.x * 16 + 9] = w_free;
    }
  } else {
    // ===================== MMA issuer ==========================================================
    constexpr uint32_t idesc = make_idesc_tf32(128, 256);   // B = [hi; lo] stacked: D = [HH | HL]
    const uint32_t hl_base = smem_u32(hl);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t ci = 0, acc_iter = 0;
    const long long t_role = clock64();
    long long w_hl = 0, w_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++acc_iter) {
      const uint32_t as = acc_iter & 1u;
      w_acc += mbar_wait(&acc_empty[as], ((acc_iter >> 1) & 1u) ^ 1u, dbg, 730);
      tcgen05_fence_after();
      const uint32_t d_acc = tmem_u + as * 256;               // columns 0..127 HH, 128..255 HL
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t u = ci % kHlStages;
        w_hl += mbar_wait(&hl_full[u], (ci / kHlStages) & 1u, dbg, 731);
        tcgen05_fence_after();
        // hi and lo tiles are adjacent: as a 256-row K-major operand they are [hi; lo]
        const uint64_t hdesc = make_kmajor_sw128_desc(hl_base + u * 2 * kChunkBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)                         // 8 fp32 (32 bytes) per K step
            umma_tf32(d_acc, hdesc + 2 * k, hdesc + 2 * k, idesc, (ch | k) != 0 ? 1u : 0u);
          umma_commit<1>(&hl_empty[u]);
          if (ch == nchunks - 1) umma_commit<1>(&acc_full[as]);
        }
        __syncwarp();
      }
    }
    if (stall != nullptr && lane == 0) {
      stall[blockIdx.x * 16 + 0] = clock64() - t_role;
      stall[blockIdx.x * 16 + 1] = w_hl;
      stall[blockIdx.x * 16 + 3] = w_acc;
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 13) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

cudaError_t launch_similarity_tc(const CUtensorMap* x_tmap, int B, int N, int D, const float* thr, int32_t* ref_ok,
                                 int32_t* medoid, float* cos, int32_t* dbg, unsigned long long* stall, int num_sms,
                                 cudaStream_t stream) {
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qc_sim_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int q = 128 / N;
  const int ntiles = (B + q - 1) / q;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  qc_sim_tc_kernel<<<grid, kTcThreads, kTcSmem, stream>>>(*x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall);
  return cudaGetLastError();
}

}  // namespace qc
