// qc_similarity_tc.cu -- K3 for pools of 9..64 candidates on the 5th-gen tensor cores  (sm_100a)
//
// The Gram matrix X X^T of a decision is GEMM-shaped and, at 2 N^2 D flop per 4 N D bytes, beyond what
// the fp32 FMA pipe (or the legacy mma.sync path, profiles/sim_sweep_r1.md) sustains against HBM.  Here:
//   * a tile is 128 consecutive rows of the [B*N x D] embedding matrix = floor(128/N) whole decisions (only
//     those rows are fetched; the rest of the MMA tile is stale shared memory that reaches discarded
//     accumulator entries only); TMA streams it in 32-column fp32 chunks (128 B rows, 128-byte swizzle)
//     through a ring of 10-11 stages of 16 KB;
//   * kind::tf32 reads an fp32 operand by ignoring its 13 low mantissa bits, so the landed tile IS the "hi"
//     operand H when it is read as B.  Eight splitter warps (two groups, alternate chunks) read each row once,
//     split it into h = trunc_tf32(x) and the exact remainder l = x - h, and store both as A operands into
//     TENSOR MEMORY (tcgen05.st; ring of 4 x 64 columns) -- nothing is written back to shared memory;
//   * one elected lane issues, per K step of 8, two tcgen05.mma kind::tf32 M = 128, N = 128 with A from TMEM
//     and B = the raw tile:  HH += H H^T  and  LH += L H^T  into two 128-column accumulators (256 TMEM columns;
//     the other 256 hold the A ring, so the accumulators are single-buffered: the epilogue pulls them into
//     registers and releases them before it does anything else).  Since HL = LH^T,
//     G = HH + (LH + LH^T) is the 3xTF32 product: fp32-accurate (<= 2e-6 from the f64 oracle even for inputs
//     living in the dropped bits; the tolerance is 1e-3);
//   * four epilogue warps read the accumulator rows back with tcgen05.ld, keep only the columns of the
//     row's own decision (the off-diagonal blocks of the 128 x 128 tile pair different decisions and are
//     discarded), and finish reciprocal norms / cosine / row mean / reference verdict / medoid vote through
//     shared memory.  Every role polls its barriers with plain mbarrier.try_wait: the roles hand chunks to each other
//     through a 4-stage TMEM ring, so the wake-up latency of a parked wait (try_wait with a suspend-time hint, which
//     the stacked-B kernel of rounds 1-2 preferred) is on the critical path here -- 670 against 720 cycles per chunk.
// Per 16 KB chunk: 8 MMAs of 128x128x8 (512 tensor cycles) and 64 KB of shared-memory traffic (TMA 16, splitters
// 16, MMA B operands 32).  Rounds 1-2 used A = H from shared memory and B = [H; L] stacked, with the lo tile written
// back to shared memory (96 KB per chunk): 0.71-0.74 of the HBM peak where this one gives 0.78-0.81
// (profiles/k3_tc_tmem_a_r2.md has the same-box comparisons, the ablations that locate what is left -- the splitters' LDS
// against the asynchronous proxy -- and the formulations that were tried and dropped).
//
// Cosine and verdict definitions are those of qc_similarity.cu (reference
// lib/quoracle/actions/consensus_rules.ex:189-210,277-286; medoid: SURVEY.md section 8(a) row A13).
#include "qc_common.cuh"
#include "ptx.cuh"

#ifndef QC_TC_ABLATE
#define QC_TC_ABLATE 0      // timing experiments only (wrong results): 1 = no L MMAs, 2 = splitters skip the row reads
#endif
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
constexpr int kTmemStages = 4;               // ring of (h, l) A operands in TMEM, 64 columns each, behind the accumulators
constexpr int kAccCols = 128;                // HH in columns 0..127, LH in 128..255, then the A ring at column 256
static_assert(kTmemStages % kSplitGroups == 0, "a TMEM stage must always belong to the same splitter group");
// Shared memory: alignment slack, kStages x 16 KB raw tiles, two [128 (+ tail)][kWin] fp32 windows (HH, LH), norms
// and row means, barriers.
// The epilogue reads up to 7 entries past a decision's last column / row (blocks of 8, results discarded); those
// reads must land in memory nobody writes, so each array they can run off carries its own tail: 7 extra window rows
// after the LH window, 16 extra entries after the norms and after the row means.
constexpr int kVecLen = 144;                 // 128 + over-read tail
// kTail = 7: the LH window carries seven never-written rows for the transposed over-reads.  kTail = 0 (pools of
// 49..64, where ten stages and 65-column windows fill the 227 KB to the last few dozen bytes): those reads are
// clamped to the decision's last row instead, and the alignment slack is 512 bytes -- the dynamic shared window of a
// kernel without static shared memory starts 1 KB into the SM's shared memory, i.e. already 1024-aligned; the
// kernel checks and bails out through the watchdog record if that ever stops being true.
__host__ __device__ constexpr int tc_slack(int tail) { return tail == 0 ? 512 : 1024; }
constexpr int tc_smem_bytes(int stages, int win, int tail) {
  return tc_slack(tail) + stages * kChunkBytes + (2 * 128 + tail) * win * 4 + 2 * kVecLen * 4 + 8 + 320;
}

// kind::tf32, fp32 accumulate, both operands K-major: c_format=1 (bit 4), a/b_format=2 (bits 7, 10), N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] B[smem]^T: the A operand is 128 lanes x 8 columns of 32-bit values at tmem_a
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
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
  uint8_t* stages = smem;                                     // kStages raw tiles of 16 KB
  float* whh = reinterpret_cast<float*>(stages + kStages * kChunkBytes);     // [128][kWin] HH window
  float* wlh = whh + 128 * kWin;                                             // [128 + tail][kWin] LH window
  float* diag = wlh + (128 + kTail) * kWin;                                  // [128 + tail] reciprocal row norms
  float* rmean = diag + kVecLen;                                             // [128 + tail] row means
  // ((256 + kTail) * kWin + 288) floats precede this point: an odd count leaves it 4-byte aligned, mbarriers need 8
  uint64_t* bars = reinterpret_cast<uint64_t*>(rmean + kVecLen + (((256 + kTail) * kWin) & 1));
  uint64_t* raw_full = bars;                                  // TMA landed the raw tile
  uint64_t* stage_empty = raw_full + kStages;                 // the MMAs that read the stage retired
  uint64_t* a_full = stage_empty + kStages;                   // splitters stored (h, 2l) of a chunk into TMEM stage t
  uint64_t* a_empty = a_full + kTmemStages;                   // the MMAs that read TMEM stage t retired
  uint64_t* acc_full = a_empty + kTmemStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q = 128 / N;                                      // whole decisions per tile
  const int ntiles = (B + q - 1) / q;
  const int nchunks = (D + 31) / 32;
  // Live rows: a tile holds q * N <= 128 rows of data, and the B operand is sized to them: UMMA N = LN = q * N rounded
  // up to 16 (the M = 128 shapes want N % 16 == 0).  Pools that do not divide 128 (48: 96 rows) then cost neither the MMA
  // cycles nor the operand reads of the rows they do not have.
  const int LN = (q * N + 15) & ~15;

  if (warp == 4 + kSplitWarps && lane == 0) {
    tma_prefetch_desc(&x_tmap);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&raw_full[s], 1);
      // tcgen05.commit of the MMAs that read the stage + one arrival per warp of every splitter group that does NOT own
      // the chunk, made after that warp has observed the chunk's raw_full phase: the producer cannot re-arm raw_full[s]
      // for chunk c + kStages before every waiter has seen chunk c's phase, so no waiter can ever find the barrier two
      // phases ahead of it (a parity wait cannot tell phase k from k - 2).  (The owning group needs no arrival of its
      // own: the MMAs whose commit is the other arrival were issued after its a_full arrival, i.e. after its reads.)
      mbar_init(&stage_empty[s], 1 + 4 * (kSplitGroups - 1));
    }
    for (int t = 0; t < kTmemStages; ++t) {
      mbar_init(&a_full[t], 4);                   // one arrival per warp of the owning splitter group
      mbar_init(&a_empty[t], 1);                  // tcgen05.commit
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_empty[0], 4);        // one arrival per epilogue warp
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
      const int b = tile * q + d;
      const bool live = row_live && b < B;
      const float th = live ? thr[b] : 0.f;                   // in flight while the tile's MMAs finish
      w_acc += mbar_wait(&acc_full[0], acc_iter & 1u, dbg, 700);
      long long tp = clock64();
      tcgen05_fence_after();
      // accumulator rows -> shared windows (only the columns of the row's own decision; 32-column groups
      // that no row of this warp needs are skipped warp-uniformly).  The accumulators are single-buffered, so the
      // next tile's MMAs wait for acc_empty: when the warp needs ONE column group (pools that divide 32, i.e. 16
      // and 32, and whenever the warp's rows fall into one 32-column group) both halves are pulled into registers
      // and the accumulator is released before the window is written.
      const int col_lo = ((warp * 32) / N) * N;
      const int col_hi = ((warp * 32 + 31) / N) * N + N;
      const int g_first = col_lo >> 5, g_last = (col_hi - 1) >> 5;
      if (g_first == g_last) {
        const int c0 = g_first << 5;
        uint32_t vh[32], vl[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + c0, vh);
        tmem_ld_32x32b_x32(tmem_base + lane_addr + LN + c0, vl);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[0]);
        if (row_live) {
          float* dh = whh + i * kWin - base;
          float* dl = wlh + i * kWin - base;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int c = c0 + k - base;
            if (c >= 0 && c < N) {
              dh[c0 + k] = __uint_as_float(vh[k]);
              dl[c0 + k] = __uint_as_float(vl[k]);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          float* dstw = (half == 0 ? whh : wlh) + i * kWin - base;
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            if (c0 + 32 <= col_lo || c0 >= col_hi) continue;
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + half * LN + c0, v);
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
        if (lane == 0) mbar_arrive(&acc_empty[0]);
      }
      named_bar_sync(1, 128);
      t_copy += clock64() - tp; tp = clock64();
      // Gram diagonal G_ii = HH_ii + 2 LH_ii -> reciprocal row norm (0 for a zero row: its cosines are 0,
      // reference consensus_rules.ex:205-209)
      if (row_live) {
        const float nrm2 = (wlh[i * kWin + r] + wlh[i * kWin + r]) + whh[i * kWin + r];
        diag[i] = nrm2 > 0.f ? rsqrtf(nrm2) : 0.f;
      }
      named_bar_sync(1, 128);
      int ok = 1;
      if (live) {
        const float ni = diag[i];
        float s = 0.f;
        // G_rj = HH_rj + (LH_rj + LH_jr).  HH = H H^T comes out of the tensor core bitwise symmetric (exact
        // tf32 products, same k order for (r, j) and (j, r)), and the bracket is symmetric by commutativity, so
        // cos(r, j) == cos(j, r) exactly (asserted in tests/test_gpu_similarity.py).
        const float* hh_row = whh + i * kWin;
        const float* lh_row = wlh + i * kWin;
        const float* lh_col = wlh + base * kWin + r;
        const float* nrm = diag + base;                       // reciprocal norms
        // cos(r, j) == cos(j, r) bitwise, so row r's values are stored down column r of the output:
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
            lr[k] = lh_row[j0 + k];
            // without tail rows the transposed read is clamped to the decision's last row (result discarded)
            lc[k] = lh_col[(kTail == 0 ? (j0 + k < N ? j0 + k : N - 1) : j0 + k) * kWin];
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
      named_bar_sync(1, 128);                                 // the windows are reused by the next tile
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
    // ===================== splitters: raw fp32 row -> (h, l) A operands in TMEM =======================
    const int p = (threadIdx.x - 128) & 127;                  // position inside the group == tile row == TMEM lane
    const int grp = (warp - 4) >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>((p >> 5) * 32) << 16;   // the warp's own TMEM lane quadrant
    const bool row_live = p < q * N;
    const long long t_role = clock64();
    long long w_raw = 0, w_tmem = 0, t_ld = 0, t_st = 0;
    uint32_t ci = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t s = ci % kStages;
        // Every group observes EVERY phase of raw_full, its own chunks or not: an mbarrier parity wait cannot tell
        // phase k from phase k - 2, so a group that skipped the chunks of the other groups could ask for phase k of a
        // stage whose phase k - 1 (another group's chunk) has not landed yet -- TMA completions are not ordered --
        // and be let through on stale data.  Observing all phases keeps every waiter within one
        // phase of the barrier; and it cannot fall two behind, because a non-owning warp's observation is itself one
        // of the arrivals the producer needs (stage_empty) before it refills the stage.
        w_raw += mbar_wait(&raw_full[s], (ci / kStages) & 1u, dbg, 710);
        if (static_cast<int>(ci % kSplitGroups) != grp) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&stage_empty[s]);        // "this warp has seen the phase" (see stage_empty's init)
          continue;
        }
        // Thread p reads its row's 32 values (eight 16-byte pieces at their swizzled places: consecutive rows hit
        // different banks) and splits each into h = what kind::tf32 will see of it and the exact remainder l, the two
        // A operands of the chunk.  The tensor core drops l's own low bits the same way: 2^-20 relative.
        const uint32_t t = ci % kTmemStages;
        const long long c0 = clock64();
        const uint8_t* row = stages + s * kChunkBytes + p * 128;
        uint32_t hi[32], lo[32];
        if (row_live && !(QC_TC_ABLATE & 2)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = *reinterpret_cast<const float4*>(row + ((j ^ (p & 7)) << 4));
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float h = tf32_trunc(xs[e]);
              hi[4 * j + e] = __float_as_uint(h);
              lo[4 * j + e] = __float_as_uint(xs[e] - h);
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) hi[e] = lo[e] = 0u;     // rows the tile does not have: accumulator rows nobody reads
        }
        const long long c1 = clock64();
        // the row is in registers before the TMEM stage is asked for: only the stores are behind the MMAs
        w_tmem += mbar_wait(&a_empty[t], ((ci / kTmemStages) & 1u) ^ 1u, dbg, 711);
        tcgen05_fence_after();
        const long long c2 = clock64();
        const uint32_t ta = tmem_base + lane_addr + 2 * kAccCols + t * 64u;
        tmem_st_32x32b_x32(ta, hi);
        tmem_st_32x32b_x32(ta + 32u, lo);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[t]);
        t_ld += c1 - c0;
        t_st += clock64() - c2;
      }
    }
    if (stall != nullptr && threadIdx.x == 128) {
      stall[blockIdx.x * 16 + 6] = clock64() - t_role;
      stall[blockIdx.x * 16 + 7] = w_raw;
      stall[blockIdx.x * 16 + 10] = w_tmem;
      stall[blockIdx.x * 16 + 14] = t_ld;
      stall[blockIdx.x * 16 + 15] = t_st;
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
        w_free += mbar_wait(&stage_empty[s], ((ci / kStages) & 1u) ^ 1u, dbg, 720);
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], static_cast<uint32_t>(q * N) * 128u);
          tma_load_2d(stages + s * kChunkBytes, &x_tmap, &raw_full[s], ch * 32, row0);
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
    const uint32_t idesc = make_idesc_tf32(128, static_cast<uint32_t>(LN));
    const uint32_t st_base = smem_u32(stages);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t ci = 0, acc_iter = 0;
    const long long t_role = clock64();
    long long w_a = 0, w_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++acc_iter) {
      w_acc += mbar_wait(&acc_empty[0], (acc_iter & 1u) ^ 1u, dbg, 730);
      tcgen05_fence_after();
      const uint32_t d_acc = tmem_u;                          // HH at column 0, LH at column LN
      for (int ch = 0; ch < nchunks; ++ch, ++ci) {
        const uint32_t u = ci % kStages;
        const uint32_t t = ci % kTmemStages;
        // The splitters' arrival follows their own wait on raw_full[u], so the raw tile (the B operand) has landed.
        w_a += mbar_wait(&a_full[t], (ci / kTmemStages) & 1u, dbg, 731);
        tcgen05_fence_after();
        const uint64_t bdesc = make_kmajor_sw128_desc(st_base + u * kChunkBytes);
        if (elect_one()) {
          const uint32_t a_h = tmem_u + 2 * kAccCols + t * 64u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                       // 8 fp32 per K step: 8 TMEM columns of A, 32 bytes of B
            umma_tf32_ts(d_acc, a_h + 8 * k, bdesc + 2 * k, idesc, (ch | k) != 0 ? 1u : 0u);             // HH (+)= H H^T
            if (!(QC_TC_ABLATE & 1))
              umma_tf32_ts(d_acc + LN, a_h + 32 + 8 * k, bdesc + 2 * k, idesc, (ch | k) != 0 ? 1u : 0u);   // LH (+)= L H^T
          }
          umma_commit<1>(&stage_empty[u]);
          umma_commit<1>(&a_empty[t]);
          if (ch == nchunks - 1) umma_commit<1>(&acc_full[0]);
        }
        __syncwarp();
      }
    }
    if (stall != nullptr && lane == 0) {
      stall[blockIdx.x * 16 + 0] = clock64() - t_role;
      stall[blockIdx.x * 16 + 1] = w_a;
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
  return launch_kernel(qc_sim_tc_kernel<kStages, kWin, kTail>, dim3(grid), dim3(kTcThreads), smem, stream, pdl, *x_tmap, B, N, D,
                       thr, ref_ok, medoid, cos, dbg, stall);
}

// Instantiations <stages, window, tail>: pools <= 32, <= 48, <= 64 (the windows grow with the pool, the ring shrinks).
#define QC_TC_VARIANTS(X) X(11, 33, 7) X(10, 49, 7) X(10, 65, 0)

// Shared-memory opt-in of the instantiations on the current device (called by qc_ctx_create).
cudaError_t similarity_tc_prepare_device() {
  cudaError_t e = cudaSuccess;
#define X(S, W, T)         \
  if (e == cudaSuccess)    \
    e = cudaFuncSetAttribute(qc_sim_tc_kernel<S, W, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(S, W, T));
  QC_TC_VARIANTS(X)
#undef X
  return e;
}

cudaError_t launch_similarity_tc(const CUtensorMap* x_tmap, int B, int N, int D, const float* thr, int32_t* ref_ok,
                                 int32_t* medoid, float* cos, int32_t* dbg, unsigned long long* stall, int num_sms,
                                 cudaStream_t stream, bool pdl) {
  if (N <= 32) return launch_tc<11, 33, 7>(x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall, num_sms, stream, pdl);
  if (N <= 48) return launch_tc<10, 49, 7>(x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall, num_sms, stream, pdl);
  return launch_tc<10, 65, 0>(x_tmap, B, N, D, thr, ref_ok, medoid, cos, dbg, stall, num_sms, stream, pdl);
}

}  // namespace qc
