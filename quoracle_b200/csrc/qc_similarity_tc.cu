// qc_similarity_tc.cu -- K3 for pools of 9..64 candidates on the 5th-gen tensor cores  (sm_100a)
//
// The Gram matrix X X^T of a decision is GEMM-shaped and, at 2 N^2 D flop per 4 N D bytes, beyond what
// the fp32 FMA pipe (or the legacy mma.sync path, profiles/sim_sweep_r1.md) sustains against HBM.  Here:
//   * a tile is 128 consecutive rows of the [B*N x D] embedding matrix = floor(128/N) whole decisions (only
//     those rows are fetched; the rest of the MMA tile is stale shared memory that reaches discarded
//     accumulator entries only); TMA streams it in 32-column fp32 chunks (128 B rows, 128-byte swizzle)
//     through a ring of 4 (pools > 32) or 5 stages, each stage = raw tile + lo tile, adjacent;
//   * kind::tf32 reads an fp32 operand by ignoring its 13 low mantissa bits, so the landed tile IS the "hi"
//     operand; four splitter warps write only the exact remainder lo = x - trunc_tf32(x) next to it (same
//     swizzled offsets) and publish it to the async proxy;
//   * one elected lane issues tcgen05.mma kind::tf32, M = 128, N = 256 (B = [raw; lo]), K = 8:
//     [HH | HL] += H [H; L]^T into a 256-column TMEM accumulator, double-buffered (512 columns).
//     G = HH + (HL + HL^T) is 3xTF32 with the symmetric term taken from the transpose: fp32-accurate
//     (<= 2e-6 from the f64 oracle even for inputs living in the dropped bits; the tolerance is 1e-3);
//   * four epilogue warps read the accumulator rows back with tcgen05.ld, keep only the columns of the
//     row's own decision (the off-diagonal blocks of the 128 x 128 tile pair different decisions and are
//     discarded), and finish reciprocal norms / cosine / row mean / reference verdict / medoid vote through
//     shared memory; waiting roles park on mbarrier.try_wait with a suspend-time hint.
// Per 16 KB chunk: 4 MMAs of 128x256x8 (512 tensor cycles) and 96 KB of shared-memory traffic (TMA 16, splitters
// 32, MMA operands 48) -- shared-memory bandwidth is what bounds the kernel (profiles/k3_tc_full_r1_summary.md).
//
// Cosine and verdict definitions are those of qc_similarity.cu (reference
// lib/quoracle/actions/consensus_rules.ex:189-210,277-286; medoid: SURVEY.md section 8(a) row A13).
#include "qc_common.cuh"
#include "ptx.cuh"

#ifndef QC_TC_SPLIT_GROUPS
#define QC_TC_SPLIT_GROUPS 2
#endif

namespace qc {

using namespace ptx;

namespace {

constexpr int kSplitGroups = QC_TC_SPLIT_GROUPS;   // splitter groups of 4 warps; chunk c belongs to group c % kSplitGroups
constexpr int kSplitWarps = 4 * kSplitGroups;
constexpr int kTcThreads = (4 + kSplitWarps + 2) * 32;   // epilogue, splitters, TMA producer, MMA issuer
constexpr int kChunkBytes = 128 * 128;       // 128 rows x 32 fp32
// Shared memory: alignment slack, kStages x (raw 16 KB + lo 16 KB), two [128][kWin] fp32 windows, norms and row
// means, barriers; five stages for every pool size (the widest windows only just fit, see kTail).
// The epilogue reads up to 7 entries past a decision's last column / row (blocks of 8, results discarded); those
// reads must land in memory nobody writes, so each array they can run off carries its own tail: 7 extra window rows
// after the HL window, 16 extra entries after the norms and after the row means.
constexpr int kVecLen = 144;                 // 128 + over-read tail
// kTail = 7: the HL window carries seven never-written rows for the transposed over-reads.  kTail = 0 (pools of
// 49..64, where five stages and 65-column windows fill the 227 KB to the last few hundred bytes): those reads are
// clamped to the decision's last row instead, and the alignment slack is 512 bytes -- the dynamic shared window of a
// kernel without static shared memory starts 1 KB into the SM's shared memory, i.e. already 1024-aligned; the
// kernel checks and bails out through the watchdog record if that ever stops being true.
__host__ __device__ constexpr int tc_slack(int tail) { return tail == 0 ? 512 : 1024; }
constexpr int tc_smem_bytes(int stages, int win, int tail) {
  return tc_slack(tail) + stages * 2 * kChunkBytes + (2 * 128 + tail) * win * 4 + 2 * kVecLen * 4 + 8 + 256;
}

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

// What kind::tf32 sees of an fp32 operand: the 13 low mantissa bits are ignored.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

}  // namespace

template <int kStages, int kWin, int kTail>
__global__ void __launch_bounds__(kTcThreads, 1)
qc_sim_tc_kernel(const __grid_constant__ CUtensorMap x_tmap,   // emb as [B*N rows x D cols] fp32, box 32 x (128/N)*N
                 int B, int N, int D, const float* __restrict__ thr, int32_t* __restrict__ ref_ok,
                 int32_t* __restrict__ medoid, float* __restrict__ cos, int32_t* __restrict__ dbg,
                 unsigned long long* __restrict__ stall) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t misalign = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (misalign > static_cast<uint32_t>(tc_slack(kTail))) {    // uniform: every thread leaves before any barrier
    if (threadIdx.x == 0 && atomicCAS(dbg, 0, 1) == 0) {
      dbg[1] = 799;
      dbg[2] = static_cast<int32_t>(blockIdx.x);
      dbg[3] = static_cast<int32_t>(misalign);
    }
    return;
  }
  uint8_t* smem = smem_raw + misalign;                        // pointer arithmetic keeps the shared address space
  uint8_t* stages = smem;                                     // kStages x (raw 16 KB, lo 16 KB)
  float* whh = reinterpret_cast<float*>(stages + kStages * 2 * kChunkBytes);   // [128][kWin] HH window
  float* whl = whh + 128 * kWin;                                             // [128][kWin] HL window
  float* diag = whl + (128 + kTail) * kWin;                                  // [128 + tail] reciprocal row norms
  float* rmean = diag + kVecLen;                                             // [128 + tail] row means
  // ((256 + kTail) * kWin + 288) floats precede this point: an odd count leaves it 4-byte aligned, mbarriers need 8
  uint64_t* bars = reinterpret_cast<uint64_t*>(rmean + kVecLen + (((256 + kTail) * kWin) & 1));
  uint64_t* raw_full = bars;                                  // TMA landed the raw tile
  uint64_t* lo_full = raw_full + kStages;                     // splitters wrote the lo tile
  uint64_t* stage_empty = lo_full + kStages;                  // the MMAs that read the stage retired
  uint64_t* acc_full = stage_empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q = 128 / N;                                      // whole decisions per tile
  const int ntiles = (B + q - 1) / q;
  const int nchunks = (D + 31) / 32;
  // Live rows: a tile holds q * N <= 128 rows of data.  The lo tile is placed right behind them (rounded up to the
  // 8-row swizzle atom, which keeps it 1024-byte aligned) and the stacked B operand is sized to what is there:
  // UMMA N = 2 * L8 instead of 256.  Pools that do not divide 128 (48: 96 rows, 24 / 40: 120) then cost neither the
  // MMA cycles nor the operand reads nor the splitter traffic of the rows they do not have.
  const int L8 = (q * N + 7) & ~7;

  if (warp == 4 + kSplitWarps && lane == 0) {
    tma_prefetch_desc(&x_tmap);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&lo_full[s], 4);                  // one arrival per warp of the owning splitter group
      // tcgen05.commit of the MMAs that read the stage + one arrival per warp of every splitter group that does NOT own
      // the chunk, made after that warp has observed the chunk's raw_full phase: the producer cannot re-arm raw_full[s]
      // for chunk c + kStages before every waiter has seen chunk c's phase, so no waiter can ever find the barrier two
      // phases ahead of it (a parity wait cannot tell phase k from k - 2).  (The owning group needs no arrival of its
      // own: its lo tile is read by the MMAs whose commit is the other arrival.)
      //
      // compute-sanitizer --tool racecheck reports the lo-tile stores below as write-after-write hazards between the
      // two groups (with five stages and two groups, consecutive laps of a stage are written by different warps).
      // The order between lap k's store and lap k + 1's is: store -> arrive(lo_full) -> MMA issue -> tcgen05.commit ->
      // stage_empty -> TMA refill -> raw_full -> the other group's wait -> store; the tool follows thread arrivals but
      // not tcgen05.commit / TMA completions, so it cannot see the chain (an experiment that added thread arrivals of
      // the owners to stage_empty did not change its report; profiles/sanitizer_r2.md).
      mbar_init(&stage_empty[s], 1 + 4 * (kSplitGroups - 1));
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);      // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 5 + kSplitWarps) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                      // K2's rows; the setup above may have run under the previous kernel's tail
  griddep_launch_dependents();

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
      w_acc += mbar_wait_parked(&acc_full[as], (acc_iter >> 1) & 1u, dbg, 700);
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
          tmem_ld_32x32b_x32(tmem_base + lane_addr + as * 256 + half * L8 + c0, v);
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
      // Gram diagonal G_ii = HH_ii + 2 HL_ii -> reciprocal row norm (0 for a zero row: its cosines are 0,
      // reference consensus_rules.ex:205-209)
      if (row_live) {
        const float nrm2 = (whl[i * kWin + r] + whl[i * kWin + r]) + whh[i * kWin + r];
        diag[i] = nrm2 > 0.f ? rsqrtf(nrm2) : 0.f;
      }
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
        const float* nrm = diag + base;                       // reciprocal norms
        // cos(r, j) == cos(j, r) bitwise (below), so row r's values are stored down column r of the output:
        // consecutive lanes (consecutive r) then write consecutive addresses
        float* cos_col = cos != nullptr ? cos + static_cast<size_t>(b) * N * N + r : nullptr;
        // Blocks of 8 columns, all loads of a block issued before use (shared memory is kept busy by the MMA
        // operand reads, so its latency is long here).  Loads past column N - 1 land in another row of the window
        // or in the arrays' own never-written tails, and their results are discarded.
        for (int j0 = 0; j0 < N; j0 += 8) {
          float up[8], lr[8], lc[8], nj[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            up[k] = hh_row[j0 + k];
            lr[k] = hl_row[j0 + k];
            // without tail rows the transposed read is clamped to the decision's last row (result discarded)
            lc[k] = hl_col[(kTail == 0 ? (j0 + k < N ? j0 + k : N - 1) : j0 + k) * kWin];
            nj[k] = nrm[j0 + k];
          }
          // branch-free so the eight dependency chains interleave
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = j0 + k;
            const bool in = j < N;
            const float g = (lr[k] + lc[k]) + up[k];
            const float c = in ? g * (ni * nj[k]) : 0.f;        // ni * nj commutes: cos(r, j) == cos(j, r) bitwise
            ok &= (!in || j == 0 || c >= th) ? 1 : 0;         // only row 0's verdict is used
            s += c;                                           // j ascending, diagonal included
            if (in && cos_col != nullptr) cos_col[j * N] = c;
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
          for (int k = 0; k < 8; ++k) m[k] = rmean[i + j0 + k];   // over-reads land in rmean's own tail
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
    // ===================== splitters: raw fp32 chunk -> lo tile (the raw tile is the hi operand) =====
    const int p = (threadIdx.x - 128) & 127;                  // position inside the group
    const int grp = (warp - 4) >> 2;
    const long long t_role = clock64();
    long long w_raw = 0;
    uint32_t ci = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t s = ci % kStages;
        // Every group observes EVERY phase of raw_full, its own chunks or not: an mbarrier parity wait cannot tell
        // phase k from phase k - 2, so a group that skipped the chunks of the other groups could ask for phase k of a
        // stage whose phase k - 1 (another group's chunk, five chunks earlier) has not landed yet -- TMA completions
        // are not ordered -- and be let through on stale data.  Observing all phases keeps every waiter within one
        // phase of the barrier; and it cannot fall two behind, because a non-owning warp's observation is itself one
        // of the arrivals the producer needs (stage_empty) before it refills the stage.
        w_raw += mbar_wait_parked(&raw_full[s], (ci / kStages) & 1u, dbg, 710);
        if (static_cast<int>(ci % kSplitGroups) != grp) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&stage_empty[s]);        // "this warp has seen the phase" (see stage_empty's init)
          continue;
        }
        const float4* src = reinterpret_cast<const float4*>(stages + s * 2 * kChunkBytes);
        float4* dl = reinterpret_cast<float4*>(stages + s * 2 * kChunkBytes + L8 * 128);   // lo tile behind the live rows
        // L8 * 8 16-byte pieces (1024 for a full tile), consecutive per warp; all loads first so their latencies overlap
        const int npieces = L8 * 8;
        float4 x[8];
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (it * 128 + p < npieces) x[it] = src[it * 128 + p];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          // the tensor core reads the raw tile as TF32 by ignoring the 13 low mantissa bits; lo is the exact
          // fp32 remainder of that truncation (its own low bits are dropped the same way: 2^-20 relative)
          if (it * 128 + p < npieces) {
            float4 l;
            l.x = x[it].x - tf32_trunc(x[it].x); l.y = x[it].y - tf32_trunc(x[it].y);
            l.z = x[it].z - tf32_trunc(x[it].z); l.w = x[it].w - tf32_trunc(x[it].w);
            dl[it * 128 + p] = l;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&lo_full[s]);
      }
    }
    if (stall != nullptr && threadIdx.x == 128) {
      stall[blockIdx.x * 16 + 6] = clock64() - t_role;
      stall[blockIdx.x * 16 + 7] = w_raw;
    }
  } else if (warp == 4 + kSplitWarps) {
    // ===================== TMA producer (warp-convergent, elected issue) =====================
    uint32_t ci = 0;
    const long long t_role = clock64();
    long long w_free = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int row0 = tile * q * N;
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t s = ci % kStages;
        w_free += mbar_wait_parked(&stage_empty[s], ((ci / kStages) & 1u) ^ 1u, dbg, 720);
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], static_cast<uint32_t>(q * N) * 128u);
          tma_load_2d(stages + s * 2 * kChunkBytes, &x_tmap, &raw_full[s], ch * 32, row0);
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
    const uint32_t idesc = make_idesc_tf32(128, 2u * static_cast<uint32_t>(L8));   // B = [hi; lo] stacked: D = [HH | HL]
    const uint32_t st_base = smem_u32(stages);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t ci = 0, acc_iter = 0;
    const long long t_role = clock64();
    long long w_hl = 0, w_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++acc_iter) {
      const uint32_t as = acc_iter & 1u;
      w_acc += mbar_wait_parked(&acc_empty[as], ((acc_iter >> 1) & 1u) ^ 1u, dbg, 730);
      tcgen05_fence_after();
      const uint32_t d_acc = tmem_u + as * 256;               // columns 0..L8-1 HH, L8..2*L8-1 HL
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t u = ci % kStages;
        w_hl += mbar_wait_parked(&lo_full[u], (ci / kStages) & 1u, dbg, 731);
        tcgen05_fence_after();
        // the lo tile starts right behind the L8 live rows of the raw tile: as a 2 * L8-row K-major operand they are [hi; lo]
        const uint64_t hdesc = make_kmajor_sw128_desc(st_base + u * 2 * kChunkBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)                         // 8 fp32 (32 bytes) per K step
            umma_tf32(d_acc, hdesc + 2 * k, hdesc + 2 * k, idesc, (ch | k) != 0 ? 1u : 0u);
          umma_commit<1>(&stage_empty[u]);
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
  if (warp == 5 + kSplitWarps) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int kStages, int kWin, int kTail>
static cudaError_t launch_tc(const CUtensorMap* x_tmap, int B, int N, int D, const float* thr, int32_t* ref_ok,
                             int32_t* medoid, float* cos, int32_t* dbg, unsigned long long* stall, int num_sms,
                             cudaStream_t stream, bool pdl) {
  constexpr int smem = tc_smem_bytes(kStages, kWin, kTail);
  static_assert(smem <= 232448, "tcgen05 K3: shared-memory carve-up exceeds the 227 KB opt-in limit");
  const int q = 128 / N;
  const int ntiles = (B + q - 1) / q;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  return launch_kernel(qc_sim_tc_kernel<kStages, kWin, kTail>, dim3(grid), dim3(kTcThreads), smem, stream, pdl, *x_tmap, B, N, D, thr,
                       ref_ok, medoid, cos, dbg, stall);
}

// Shared-memory opt-in of the three instantiations on the current device (called by qc_ctx_create).
cudaError_t similarity_tc_prepare_device() {
  cudaError_t e = cudaFuncSetAttribute(qc_sim_tc_kernel<5, 33, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(5, 33, 7));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(qc_sim_tc_kernel<5, 49, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(5, 49, 7));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(qc_sim_tc_kernel<5, 65, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(5, 65, 0));
}

cudaError_t launch_similarity_tc(const CUtensorMap* x_tmap, int B, int N, int D, const float* thr, int32_t* ref_ok,
                                 int32_t* medoid, float* cos, int32_t* dbg, unsigned long long* stall, int num_sms,
                                 cudaStream_t stream, bool pdl) {
  if (N <= 32) return launch_tc<5, 33, 7>(x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall, num_sms, stream, pdl);
  if (N <= 48) return launch_tc<5, 49, 7>(x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall, num_sms, stream, pdl);
  return launch_tc<5, 65, 0>(x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall, num_sms, stream, pdl);
}

}  // namespace qc
