// qc_encoder.cu -- K1: token gather + tcgen05 GEMM + GELU + per-block token sum  (sm_100a)
//
// Computes, for every candidate row r and every block j of TB consecutive token slots,
//   part[r, j, d] = sum_{t in block j, t < len[r]} GELU_tanh( sum_k E[tok[r,t], k] * W[k, d] )
// K2 (qc_similarity.cu: qc_finalize_kernel) adds the blocks in a fixed order, divides by len and
// L2-normalises.  This is the local stand-in for the provider call behind
// Embeddings.get_embedding/2 (reference lib/quoracle/models/embeddings.ex:277-298); the encoder
// definition is this build's, the reference fixes none.
//
// Tiling (DESIGN.md "K1"): the GEMM is issued TRANSPOSED -- M = output features (TMEM lanes),
// N = tokens (TMEM columns) -- so that the token mean is a per-thread sum over the columns a
// thread reads back from TMEM, with no cross-lane reduction.  Per CTA:
//   * the TB x D bf16 token rows of one block are gathered from the embedding table with
//     cp.async (16 B chunks, written in the 128-byte-swizzled K-major layout tcgen05 expects)
//     and stay RESIDENT in shared memory: KB = D/64 k-blocks of TB x 128 B;
//   * W^T tiles (128 features x WBK k) stream through a W_STAGES-deep TMA ring (WBK = 64 with
//     128-byte swizzle, or 32 with 64-byte swizzle for a finer-grained, deeper ring);
//   * D/128 (or D/256 for a CTA pair) feature tiles reuse the resident tokens; the k-block
//     buffers are released one by one during the last feature tile so the next block's gather
//     overlaps the tail of this one;
//   * accumulators live in TMEM (2 stages), read back by 4 epilogue warps with tcgen05.ld.
// CG == 2 runs a CTA pair (cta_group::2): M = 256 features split across the pair, N = 2*TB
// tokens (each CTA's resident block is half of the B operand), one MMA issuer in the leader.
//
// Shared-memory bandwidth is the co-limiter (ncu: l1tex tc wavefronts == tensor active): each
// K=16 MMA reads (M + N) x 32 B of operands, and streamed W bytes are also WRITTEN once, so the
// pair variant (half the W bytes per CTA per flop) is the fast one.
//
// Warp roles (320 threads): 0-3 epilogue (TMEM lane quadrant = warp id), 4-7 gather,
// 8 W-tile TMA producer, 9 MMA issuer + TMEM allocator.
#include "qc_common.cuh"
#include "ptx.cuh"

namespace qc {

using namespace ptx;

template <int D_, int TB_, int CG_, int WBK_, int GM_, int MC_ = 1>
struct EncTraits {
  static constexpr int D = D_;
  static constexpr int TB = TB_;
  static constexpr int CG = CG_;
  static constexpr int MC = MC_;                  // CTA pairs per cluster sharing every W^T tile through TMA multicast
  static constexpr int CLUSTER = CG_ * MC_;       // CTAs per cluster
  static constexpr int WBK = WBK_;                // k elements per streamed W stage
  static constexpr int GM = GM_;                  // gather mode: 0 = cp.async (LDGSTS), 1 = TMA gather4
  static constexpr int BK = 64;                   // bf16 elements per 128-byte swizzle row (token tiles)
  // K split: when a token block does not fit next to a W ring (dim 1536 with more than 64 tokens per CTA) the
  // block's rows are held one K HALF at a time and re-gathered for every group of FG feature tiles, whose FG
  // accumulators (all of TMEM) stay live across the halves.  Twice the tokens per streamed W byte: the kernel is
  // bound by bytes delivered INTO the SM per flop, and W is two thirds of them.
  static constexpr int KS = (D * TB * 2 > 196608) ? 2 : 1;   // K parts held in turn
  static constexpr int KB = D / BK / KS;          // resident token k-blocks (of the current K part)
  static constexpr int NWS = D / WBK;             // W stages per feature tile
  static constexpr int NWSK = NWS / KS;           // ... per feature tile and K part
  static constexpr int MT = 128 * CG;             // features per MMA tile over the CTA group
  static constexpr int NFT = D / MT;              // feature tiles
  static constexpr int FG = KS > 1 ? 2 : 1;       // feature tiles accumulating side by side (one TMEM stage each)
  static constexpr int NG = NFT / FG;             // groups of them per token block
  static constexpr int NTP = KS > 1 ? NG * KS : 1;   // token passes (gathers of the resident rows) per token block
  static constexpr int UMMA_N = TB * CG;          // tokens per MMA tile
  static constexpr int TOK_KB_BYTES = TB * 128;
  static constexpr int TOK_BYTES = KB * TOK_KB_BYTES;
  static constexpr int TOK_BYTES_PAD = (TOK_BYTES + 1023) & ~1023;
  static constexpr int W_STAGE_BYTES = 128 * WBK * 2;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int SMEM_MAX = 232448;         // 227 KB opt-in limit
  static constexpr int W_STAGES_FIT = (SMEM_MAX - 1024 - BAR_BYTES - TOK_BYTES_PAD) / W_STAGE_BYTES;
  static constexpr int W_STAGES = W_STAGES_FIT > 8 ? 8 : W_STAGES_FIT;
  static constexpr int ACC_STAGES = 2;
  static constexpr int ACC_COLS = UMMA_N <= 64 ? 64 : (UMMA_N <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;   // 128 / 256 / 512: powers of two >= 32
  static constexpr int SMEM_BYTES = 1024 + TOK_BYTES_PAD + W_STAGES * W_STAGE_BYTES + BAR_BYTES;
  static constexpr int GATHER_GROUP = (KB % 4 == 0) ? 4 : 3;   // k-blocks per cp.async group / proxy fence
  static constexpr int GATHER_LANES = TB > 128 ? 16 : 8;      // gather4 issuing lanes per gather warp (4 rows each)
  static_assert(MC == 1 || (CG == 2 && (MC == 2 || MC == 4)), "W multicast needs CTA pairs; cluster of 4 or 8");
  static_assert(4 * GATHER_LANES * 4 >= TB, "gather lanes must cover the token block");
  static_assert(KS == 1 || (GM == 1 && WBK == 64 && NFT % 2 == 0 && D % (BK * KS) == 0),
                "the K-split walk needs TMA gather, 64-wide W stages and an even number of feature tiles");
  static_assert(WBK == 64 || WBK == 32, "W stage width is 64 (SW128) or 32 (SW64)");
  static_assert(D % MT == 0, "feature tiles must cover D");
  static_assert(D % BK == 0, "k-blocks must cover D");
  static_assert(TB % 8 == 0 && UMMA_N % 16 == 0 && UMMA_N <= 256, "invalid token block");
  static_assert(W_STAGES >= 2, "need at least two W stages");
  static_assert(KB % GATHER_GROUP == 0 && KB / GATHER_GROUP >= 2, "gather groups must tile the k-blocks");
  static_assert((2 * W_STAGES + 2 * KB + 2 * ACC_STAGES) * 8 + 8 <= BAR_BYTES, "barrier region too small");
};

constexpr int kEncThreads = 320;

// Sum over tokens of GELU_tanh(x) = 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) )), accumulated as
// 0.5 (sum x + sum x tanh(.)): five FP32 operations and one MUFU per element instead of six and one (the epilogue is
// the busiest non-tensor role; at dim 384 it is the bound).
struct GeluSum {
  float s1 = 0.f, s2 = 0.f;
  __device__ __forceinline__ void add(float x) {
    const float x2 = x * x;
    const float inner = x * fmaf(0.0356774081f, x2, 0.7978845608f);
    s2 = fmaf(x, tanh_approx(inner), s2);
    s1 += x;
  }
  __device__ __forceinline__ float value() const { return 0.5f * (s1 + s2); }
};

__device__ __forceinline__ int block_valid(const int32_t* __restrict__ lens, int blk, int nblocks, int bpc, int TB) {
  if (blk >= nblocks) return 0;
  if (lens == nullptr) return TB;                 // packed token stream: every block is full (padding ids are -1)
  const int cand = blk / bpc;
  const int j = blk - cand * bpc;
  int v = lens[cand] - j * TB;
  v = v < 0 ? 0 : v;
  return v > TB ? TB : v;
}

// Work slot -> token block.  Dense mode (map == nullptr): slot == block, blocks past a candidate's length are
// skipped by every role.  Compacted mode: `map` lists the non-empty blocks (built by the host entry points from
// the lengths they already hold), so no CTA pair runs half empty and no item is skipped -- ragged batches then
// cost what their non-empty blocks cost.  A slot past the end maps to block `nblocks`, which block_valid() rejects.
__device__ __forceinline__ int slot_block(const int32_t* __restrict__ map, int slot, int nslots, int nblocks) {
  if (slot >= nslots) return nblocks;
  return map != nullptr ? map[slot] : slot;
}

// Sum of GELU over NCOL consecutive TMEM columns of this thread's lane, software-pipelined:
// the next 32-column load is in flight while the current one is consumed.
template <int NCOL>
__device__ __forceinline__ float gelu_sum_columns(uint32_t taddr) {
  static_assert(NCOL % 8 == 0, "column count must be a multiple of 8");
  constexpr int FULL = NCOL / 32;
  constexpr int REM = NCOL % 32;
  GeluSum sum;
  uint32_t va[32], vb[32];
  if constexpr (FULL > 0) tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
  for (int c = 0; c < FULL; ++c) {
    tmem_ld_wait();
    if (c + 1 < FULL) {
      if (c & 1) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, va); else tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) sum.add(__uint_as_float((c & 1) ? vb[i] : va[i]));
  }
  if constexpr (REM >= 16) {
    uint32_t vt[16];
    tmem_ld_32x32b_x16(taddr + FULL * 32, vt);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) sum.add(__uint_as_float(vt[i]));
  }
  if constexpr (REM % 16 == 8) {
    uint32_t vt[8];
    tmem_ld_32x32b_x8(taddr + FULL * 32 + (REM >= 16 ? 16 : 0), vt);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) sum.add(__uint_as_float(vt[i]));
  }
  return sum.value();
}

// Same traversal, one sum per chunk of CH columns: the packed token stream gives every CH-token chunk of a block to
// one candidate, so chunk sums are what K2 adds up.  CH is the widest of 32 / 16 / 8 that divides the token block.
__host__ __device__ constexpr int chunk_tokens(int tb) { return tb % 32 == 0 ? 32 : (tb % 16 == 0 ? 16 : 8); }

template <int NCOL, int CH>
__device__ __forceinline__ void gelu_sum_chunks(uint32_t taddr, float (&out)[NCOL / CH]) {
  static_assert(NCOL % CH == 0 && (CH == 32 || CH == 16 || CH == 8), "chunk width must divide the token block");
  constexpr int NCHUNK = NCOL / CH;
  uint32_t va[CH], vb[CH];
  auto load = [](uint32_t addr, uint32_t (&v)[CH]) {
    if constexpr (CH == 32) tmem_ld_32x32b_x32(addr, v);
    else if constexpr (CH == 16) tmem_ld_32x32b_x16(addr, v);
    else tmem_ld_32x32b_x8(addr, v);
  };
  load(taddr, va);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    tmem_ld_wait();
    if (c + 1 < NCHUNK) {
      if (c & 1) load(taddr + (c + 1) * CH, va); else load(taddr + (c + 1) * CH, vb);
    }
    GeluSum sum;
#pragma unroll
    for (int i = 0; i < CH; ++i) sum.add(__uint_as_float((c & 1) ? vb[i] : va[i]));
    out[c] = sum.value();
  }
}

template <class Tr>
__global__ void __launch_bounds__(kEncThreads, 1)
qc_encode_kernel(const __grid_constant__ CUtensorMap w_tmap,          // W^T [D x D] bf16, box WBK x 128
                 const __grid_constant__ CUtensorMap e_tmap,          // E [vocab x D] bf16, box 64 x 1 (gather4)
                 const uint16_t* __restrict__ table,                   // E [vocab x D] bf16
                 const int32_t* __restrict__ token_ids,                // [R x T]
                 const int32_t* __restrict__ lens,                     // [R]
                 float* __restrict__ part,                             // [R x bpc x D]
                 int32_t* __restrict__ dbg,                            // [0..3] watchdog, [4] token-range flag
                 unsigned long long* __restrict__ stall,               // optional [grid x 16] per-role stall cycles
                 const int32_t* __restrict__ block_map,                // optional compacted block list [nslots]
                 int nslots,
                 const int32_t* __restrict__ chunk_row,                // packed stream: partial-row id per chunk (-1 = padding)
                 int chunked,                                          // chunked: part rows come from chunk_row
                 int R, int T, int bpc, int vocab) {
  constexpr int D = Tr::D, TB = Tr::TB, CG = Tr::CG, KB = Tr::KB, NFT = Tr::NFT, WBK = Tr::WBK, NWS = Tr::NWS;
  constexpr int UMMA_N = Tr::UMMA_N, WS = Tr::W_STAGES, ACC_COLS = Tr::ACC_COLS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tok_smem = smem;                                  // KB x (TB x 128 B)
  uint8_t* w_smem = smem + Tr::TOK_BYTES_PAD;                // WS x W_STAGE_BYTES, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + WS * Tr::W_STAGE_BYTES);
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + WS;
  uint64_t* tok_full = w_empty + WS;
  uint64_t* tok_empty = tok_full + KB;
  uint64_t* acc_full = tok_empty + KB;
  uint64_t* acc_empty = acc_full + Tr::ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Tr::ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int MC = Tr::MC, CL = Tr::CLUSTER;
  const uint32_t crank = (CL > 1) ? (blockIdx.x % CL) : 0u;    // rank in the cluster (cluster dim x = CL, 1-D grid)
  const uint32_t rank = (CG == 2) ? (crank & 1u) : 0u;         // position in the CTA pair
  const uint32_t lcta = crank & ~1u;                            // cluster rank of this pair's leader (even) CTA
  const uint32_t pair = crank >> 1;                             // which pair of the cluster (0 when MC == 1)
  const bool leader = (rank == 0);
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << lcta);   // commit targets: both CTAs of this pair
  // Work: a cluster item is MC pair items (CG token blocks each).  The MC pairs of a cluster walk the same W^T
  // tiles in lock step (every tile is loaded from L2 once and multicast to all of them), so every role of every
  // CTA of the cluster takes part in a cluster item as soon as ANY of its MC * CG blocks is non-empty; a pair whose
  // own blocks are empty then multiplies zero rows and stores nothing.
  const int cl = blockIdx.x / CL;
  const int ncl = gridDim.x / CL;
  const int nblocks = R * bpc;
  const int nitems = (nslots + CG - 1) / CG;                   // pair items: CG slots (blocks) each
  const int ncitems = (nitems + MC - 1) / MC;                  // cluster items
  auto item_any = [&](int citem) -> bool {
    bool any = false;
#pragma unroll
    for (int s = 0; s < MC * CG; ++s)
      any |= block_valid(lens, slot_block(block_map, citem * MC * CG + s, nslots, nblocks), nblocks, bpc, TB) > 0;
    return any;
  };

  // ---- one-time setup -----------------------------------------------------------------
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&w_tmap);
    if (Tr::GM == 1) tma_prefetch_desc(&e_tmap);
    for (int s = 0; s < WS; ++s) {
      mbar_init(&w_full[s], 1);   // leader's producer arrive.expect_tx (TMA bytes of the whole group)
      // tcgen05.commit of every pair that reads the stage.  With multicast only pair 0's CTAs produce: their w_empty
      // collects one commit per pair; the leader of every other pair keeps its own copy (its pair's commit only) to
      // know when it may arm w_full for the next phase.
      mbar_init(&w_empty[s], crank < 2 ? MC : 1);
    }
    for (int k = 0; k < KB; ++k) {
      // cp.async gather: one arrival per gather warp of each CTA; TMA gather: the leader's expect_tx
      mbar_init(&tok_full[k], Tr::GM == 0 ? 4 * CG : 1);
      mbar_init(&tok_empty[k], 1);      // tcgen05.commit
    }
    for (int a = 0; a < Tr::ACC_STAGES; ++a) {
      mbar_init(&acc_full[a], 1);        // tcgen05.commit
      mbar_init(&acc_empty[a], 4 * CG);  // one arrival per epilogue warp of each CTA in the group
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 9) {
    tmem_alloc<CG>(tmem_slot, Tr::TMEM_COLS);
  }
  tcgen05_fence_before();
  if constexpr (CL > 1) cluster_sync(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may have run
  // under the tail of the previous kernel of the stream.  From here on this kernel reads token ids a pack kernel may
  // have produced and overwrites the partial sums the previous step's K2 read, so every thread waits for the
  // previous kernel to finish; the kernel after this one (K2) may be staged right away -- it waits the same way.
  griddep_wait();
  griddep_launch_dependents();

  if (warp < 4) {
    // ===================== epilogue: TMEM -> GELU -> token sum -> part ======================
    const int row = warp * 32 + lane;                         // feature row inside this CTA's tile half
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t acc_iter = 0;
    long long st_wait = 0;
    const long long st_t0 = clock64();
    for (int citem = cl; citem < ncitems; citem += ncl) {
      if (!item_any(citem)) continue;
      const int item = citem * MC + static_cast<int>(pair);
      int valid[CG];
#pragma unroll
      for (int s = 0; s < CG; ++s)
        valid[s] = block_valid(lens, slot_block(block_map, item * CG + s, nslots, nblocks), nblocks, bpc, TB);
      for (int ft = 0; ft < NFT; ++ft, ++acc_iter) {
        const uint32_t as = acc_iter & 1u;
        st_wait += mbar_wait(&acc_full[as], (acc_iter >> 1) & 1u, dbg, 100 + ft);
        tcgen05_fence_after();
        if (chunked) {
          constexpr int CH = chunk_tokens(TB);
          constexpr int NCH = TB / CH;
          float csum[CG][NCH];
#pragma unroll
          for (int s = 0; s < CG; ++s) gelu_sum_chunks<TB, CH>(tmem_base + lane_addr + as * ACC_COLS + s * TB, csum[s]);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(&acc_empty[as], lcta); else mbar_arrive(&acc_empty[as]);
          }
          // consecutive chunks of one candidate inside a block share a partial row (chunk_row, warp-uniform)
#pragma unroll
          for (int s = 0; s < CG; ++s) {
            if (valid[s] > 0) {
              const size_t blk = static_cast<size_t>(slot_block(block_map, item * CG + s, nslots, nblocks));
              const int32_t* cr = chunk_row + blk * NCH;
              const size_t col = static_cast<size_t>(ft) * Tr::MT + rank * 128 + row;
              int prev = cr[0];
              float acc = csum[s][0];
#pragma unroll
              for (int k = 1; k < NCH; ++k) {
                const int rk = cr[k];
                if (rk == prev) {
                  acc += csum[s][k];
                } else {
                  if (prev >= 0) part[static_cast<size_t>(prev) * D + col] = acc;
                  prev = rk;
                  acc = csum[s][k];
                }
              }
              if (prev >= 0) part[static_cast<size_t>(prev) * D + col] = acc;
            }
          }
          continue;
        }
        float sums[CG];
#pragma unroll
        for (int s = 0; s < CG; ++s) sums[s] = gelu_sum_columns<TB>(tmem_base + lane_addr + as * ACC_COLS + s * TB);
        // the accumulator stage is free as soon as its columns are in registers
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(&acc_empty[as], lcta); else mbar_arrive(&acc_empty[as]);
        }
#pragma unroll
        for (int s = 0; s < CG; ++s) {
          if (valid[s] > 0) {
            const size_t blk = static_cast<size_t>(slot_block(block_map, item * CG + s, nslots, nblocks));
            part[blk * D + ft * Tr::MT + rank * 128 + row] = sums[s];
          }
        }
      }
    }
    if (stall != nullptr && threadIdx.x == 0) {
      stall[blockIdx.x * 16 + 4] = static_cast<unsigned long long>(clock64() - st_t0);
      stall[blockIdx.x * 16 + 5] = static_cast<unsigned long long>(st_wait);
    }
  } else if (warp < 8 && Tr::GM == 1) {
    // ===================== gather (TMA gather4) ===================================================
    // Row group rg = (warp-4)*8 + lane (lanes 0..7 of each gather warp) owns rows 4rg..4rg+3: eight
    // gather4 issues per warp per k-block (the per-lane operands force the compiler's one-lane-at-a-time
    // issue loop, so the work is spread over the four gather warps).
    {
      constexpr int GL = Tr::GATHER_LANES;
      const int rg = (warp - 4) * GL + lane;
      const bool active = lane < GL && rg * 4 < TB;
      const uint32_t tok_base = smem_u32(tok_smem);
      uint32_t it = 0;
      long long st_wait = 0;
      const long long st_t0 = clock64();
      for (int citem = cl; citem < ncitems; citem += ncl) {
        if (!item_any(citem)) continue;
        const int item = citem * MC + static_cast<int>(pair);
        const int blk = slot_block(block_map, item * CG + static_cast<int>(rank), nslots, nblocks);
        const int valid = block_valid(lens, blk, nblocks, bpc, TB);
        const int cand = blk < nblocks ? blk / bpc : 0;
        const int j = blk < nblocks ? blk - cand * bpc : 0;
        const int32_t* ids = token_ids + static_cast<size_t>(cand) * T + static_cast<size_t>(j) * TB;
        int32_t rid[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rowi = rg * 4 + q;
          int32_t id = -1;
          if (active && rowi < valid) id = ids[rowi];
          if (id >= vocab) {
            atomicExch(dbg + 4, 1);
            id = -1;
          }
          rid[q] = id >= 0 ? id : vocab;          // out-of-range row -> TMA zero fill
        }
        // one pass per token block, or (K split) one per feature-tile group and K part: the same rows, the columns
        // of K part tp % KS
#pragma unroll 1
        for (int tp = 0; tp < Tr::NTP; ++tp, ++it) {
          const uint32_t par_empty = (it & 1u) ^ 1u;
          const int col0 = (tp % Tr::KS) * KB * 64;
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb) {
            // compute-sanitizer --tool racecheck reports this arrive.expect_tx against the MMA warp's try_wait on the
            // same mbarrier word (profiles/sanitizer_r*.md).  Both are mbarrier operations -- atomic on the 64-bit
            // barrier state by definition -- which the tool does not model as synchronisation; no plain load or store
            // ever touches the word after fence_barrier_init().  The data the barrier orders (TMA-written rows read by
            // tcgen05.mma) is async-proxy traffic the tool does not see at all.  The ordering argument: phase k of
            // tok_full[kb] needs this arm AND all CG * TB * 128 bytes; the arm for phase k + 1 is issued only after
            // tok_empty[kb]'s phase k (a tcgen05.commit after the last MMA that read the rows), which the waiter of
            // tok_full's phase k has necessarily passed.
            if (lane == 0) {
              st_wait += mbar_wait(&tok_empty[kb], par_empty, dbg, 200 + kb);
              if (warp == 4 && leader) mbar_expect_tx(&tok_full[kb], CG * Tr::TOK_KB_BYTES);
            }
            __syncwarp();
            if (active)
              tma_gather4<CG>(tok_base + kb * Tr::TOK_KB_BYTES + rg * 512, &e_tmap, &tok_full[kb], col0 + kb * 64, rid[0],
                              rid[1], rid[2], rid[3], lcta);
          }
        }
      }
      if (stall != nullptr && threadIdx.x == 128) {
        stall[blockIdx.x * 16 + 6] = static_cast<unsigned long long>(clock64() - st_t0);
        stall[blockIdx.x * 16 + 7] = static_cast<unsigned long long>(st_wait);
      }
    }
  } else if (warp < 8) {
    // ===================== gather: table rows -> swizzled resident token tiles ==============
    const int p = threadIdx.x - 128;           // 0..127
    const int r16 = p >> 3;                    // row within a 16-row slab
    const int c = p & 7;                       // 16-byte chunk within the 128-byte row
    const uint32_t tok_base = smem_u32(tok_smem);
    // swizzle: chunk index XOR (row mod 8); rows of one thread are congruent mod 16 so it is constant
    const uint32_t dst_off0 = static_cast<uint32_t>(r16 * 128 + ((c ^ (r16 & 7)) << 4));
    uint32_t it = 0;
    long long st_wait = 0, st_ld = 0, st_fence = 0;
    const long long st_t0 = clock64();
    for (int citem = cl; citem < ncitems; citem += ncl) {
      if (!item_any(citem)) continue;
      const int item = citem * MC + static_cast<int>(pair);
      const int blk = slot_block(block_map, item * CG + static_cast<int>(rank), nslots, nblocks);
      const int valid = block_valid(lens, blk, nblocks, bpc, TB);
      const int cand = blk < nblocks ? blk / bpc : 0;
      const int j = blk < nblocks ? blk - cand * bpc : 0;
      const int32_t* ids = token_ids + static_cast<size_t>(cand) * T + static_cast<size_t>(j) * TB;
      constexpr int QN = (TB + 15) / 16;
      const uint16_t* src[QN];
      uint32_t nbytes[QN];
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        const int rowi = q * 16 + r16;
        int32_t id = -1;
        if (rowi < valid) id = ids[rowi];
        if (id >= vocab) {
          atomicExch(dbg + 4, 1);
          id = -1;
        }
        nbytes[q] = id >= 0 ? 16u : 0u;
        src[q] = table + (id >= 0 ? static_cast<size_t>(id) * D : 0) + c * 8;
      }
      const uint32_t par_empty = (it & 1u) ^ 1u;
      // k-blocks are fetched in groups of GG: one cp.async group, one generic->async proxy fence and one
      // round of barrier arrivals per group (the fence is the expensive part: it drains behind the MMA and
      // TMA traffic of the async proxy, so one per k-block throttled the gather).
      constexpr int GG = Tr::GATHER_GROUP;
      constexpr int NG = KB / GG;
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int kk = 0; kk < GG; ++kk) {
          const int kb = g * GG + kk;
          st_wait += mbar_wait(&tok_empty[kb], par_empty, dbg, 200 + kb);
          const uint32_t dst = tok_base + kb * Tr::TOK_KB_BYTES + dst_off0;
#pragma unroll
          for (int q = 0; q < QN; ++q) {
            if (q * 16 + r16 < TB) cp_async_16(dst + q * 2048, src[q] + kb * 64, nbytes[q]);
          }
        }
        cp_async_commit();
        if (g >= 1) {
          const long long tl = clock64();
          cp_async_wait<1>();
          const long long tf = clock64();
          fence_proxy_async_smem();
          st_ld += tf - tl;
          st_fence += clock64() - tf;
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int kk = 0; kk < GG; ++kk) {
              const int kb = (g - 1) * GG + kk;
              if constexpr (CG == 2) mbar_arrive_cluster(&tok_full[kb], lcta); else mbar_arrive(&tok_full[kb]);
            }
          }
        }
      }
      {
        const long long tl = clock64();
        cp_async_wait<0>();
        const long long tf = clock64();
        fence_proxy_async_smem();
        st_ld += tf - tl;
        st_fence += clock64() - tf;
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < GG; ++kk) {
          const int kb = (NG - 1) * GG + kk;
          if constexpr (CG == 2) mbar_arrive_cluster(&tok_full[kb], lcta); else mbar_arrive(&tok_full[kb]);
        }
      }
      ++it;
    }
    if (stall != nullptr && threadIdx.x == 128) {
      stall[blockIdx.x * 16 + 6] = static_cast<unsigned long long>(clock64() - st_t0);
      stall[blockIdx.x * 16 + 7] = static_cast<unsigned long long>(st_wait);
      stall[blockIdx.x * 16 + 10] = static_cast<unsigned long long>(st_ld);
      stall[blockIdx.x * 16 + 13] = static_cast<unsigned long long>(st_fence);
    }
  } else if (warp == 8) {
    // ===================== W^T tile producer (TMA), warp-convergent like the MMA issuer ============
    {
      uint32_t wi = 0;
      long long st_wait = 0;
      const long long st_t0 = clock64();
      // With multicast (MC > 1) only pair 0 loads: CTA 0 / CTA 1 each fetch their 128-feature half of the tile ONCE and
      // multicast it to the same-parity CTA of all pairs; their w_empty[s] counts one tcgen05.commit per pair, so a
      // stage is refilled only when every pair is done with it.  Every pair's leader arms its OWN w_full[s] (one local
      // arrive.expect_tx per phase, after its pair's commit released the stage): remote arming from CTA 0 put a
      // cluster-scope release on the producer's critical path for every pair and made the multicast kernel 4x slower
      // than the unicast one.  A tile may land in a pair before that pair has armed the phase -- the barrier's
      // transaction count goes transiently negative, which is legal, and the phase cannot complete without the arm.
      for (int citem = cl; citem < ncitems && (pair == 0 || leader); citem += ncl) {
        if (!__any_sync(0xffffffffu, item_any(citem))) continue;
        // tile order = the MMA issuer's: groups of FG feature tiles; per group the K parts; per K part the FG tiles
        // (without a K split: feature tile after feature tile)
        for (int g = 0; g < Tr::NG; ++g)
        for (int kh = 0; kh < Tr::KS; ++kh)
        for (int f = 0; f < Tr::FG; ++f) {
          const int ft = g * Tr::FG + f;
          for (int wsl = 0; wsl < Tr::NWSK; ++wsl, ++wi) {
            const int ws = kh * Tr::NWSK + wsl;                   // k position of the tile inside the feature tile
            const uint32_t s = wi % WS;
            const uint32_t ph = (wi / WS) & 1u;
            st_wait += mbar_wait(&w_empty[s], ph ^ 1u, dbg, 300 + static_cast<int>(s));
            void* dst = w_smem + s * Tr::W_STAGE_BYTES;
            if (elect_one()) {
              if constexpr (MC > 1) {
                constexpr uint16_t even_mask = MC == 2 ? 0x5 : 0x55;
                if (leader) mbar_expect_tx(&w_full[s], 2 * Tr::W_STAGE_BYTES);
                if (pair == 0)
                  tma_load_2d_pair_mcast(dst, &w_tmap, &w_full[s], ws * WBK, ft * Tr::MT + static_cast<int>(rank) * 128,
                                         static_cast<uint16_t>(even_mask << rank));
              } else if constexpr (CG == 2) {
                if (leader) mbar_expect_tx(&w_full[s], 2 * Tr::W_STAGE_BYTES);
                tma_load_2d_pair(dst, &w_tmap, &w_full[s], ws * WBK, ft * Tr::MT + static_cast<int>(rank) * 128);
              } else {
                mbar_expect_tx(&w_full[s], Tr::W_STAGE_BYTES);
                tma_load_2d(dst, &w_tmap, &w_full[s], ws * WBK, ft * Tr::MT);
              }
            }
            __syncwarp();
          }
        }
      }
      if (stall != nullptr && lane == 0) {
        stall[blockIdx.x * 16 + 8] = static_cast<unsigned long long>(clock64() - st_t0);
        stall[blockIdx.x * 16 + 9] = static_cast<unsigned long long>(st_wait);
      }
    }
  } else {
    // ===================== MMA issuer (leader CTA) ===========================================
    // The whole warp runs the loop in convergent control flow (barrier waits are warp-wide); only the
    // tcgen05.mma / tcgen05.commit instructions sit under elect_one().  Issuing from a divergent
    // `lane == 0` branch makes the compiler wrap every UTCHMMA in an ELECT/R2UR waterfall loop
    // (~70 cycles per MMA), which left the tensor pipe issue-starved.
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(Tr::MT, UMMA_N);
      constexpr int WPK = 64 / WBK;                              // W stages per token k-block
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t tok_base = smem_u32(tok_smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t wi = 0, acc_iter = 0, it = 0;
      long long st_w = 0, st_tok = 0, st_acc = 0;
      const long long st_t0 = clock64();
      for (int citem = cl; citem < ncitems; citem += ncl) {
        if (!__any_sync(0xffffffffu, item_any(citem))) continue;
        // Walk: groups of FG feature tiles; per group the K parts; per K part the FG tiles of the group, each into its
        // own accumulator stage (accumulating across the K parts).  Without a K split FG = KS = 1 and this is "feature
        // tile after feature tile over resident tokens".  With it, accumulator f of a group is complete after the LAST
        // K part's f-th tile, i.e. the epilogue of tile 0 runs under the MMAs of tile 1's last part, and the epilogue
        // of tile 1 under the next group's first tile: the two TMEM stages still overlap MMA and epilogue.
        for (int g = 0; g < Tr::NG; ++g, acc_iter += Tr::FG)
        for (int kh = 0; kh < Tr::KS; ++kh, it += (Tr::KS > 1 ? 1u : 0u))
        for (int f = 0; f < Tr::FG; ++f) {
          const int ft = g * Tr::FG + f;
          const uint32_t ai = acc_iter + static_cast<uint32_t>(f);
          const uint32_t as = ai & 1u;
          if (kh == 0) {                                          // first MMAs into this accumulator: it must be drained
            if constexpr (CG == 2) st_acc += mbar_wait_cluster(&acc_empty[as], ((ai >> 1) & 1u) ^ 1u, dbg, 400 + ft);
            else st_acc += mbar_wait(&acc_empty[as], ((ai >> 1) & 1u) ^ 1u, dbg, 400 + ft);
            tcgen05_fence_after();
          }
          const uint32_t tmem_d = tmem_u + as * ACC_COLS;
          const bool tok_first = Tr::KS > 1 ? (f == 0) : (g == 0);             // first / last use of the resident rows
          const bool tok_last = Tr::KS > 1 ? (f == Tr::FG - 1) : (g == Tr::NG - 1);
          for (int ws = 0; ws < Tr::NWSK; ++ws, ++wi) {
            const int kb = ws / WPK;                              // resident token k-block this W stage multiplies
            const int sub = ws % WPK;                             // which WBK-wide part of it
            if (tok_first && sub == 0) {
              if constexpr (CG == 2) st_tok += mbar_wait_cluster(&tok_full[kb], it & 1u, dbg, 500 + kb);
              else st_tok += mbar_wait(&tok_full[kb], it & 1u, dbg, 500 + kb);
            }
            const uint32_t s = wi % WS;
            st_w += mbar_wait(&w_full[s], (wi / WS) & 1u, dbg, 600 + static_cast<int>(s));
            tcgen05_fence_after();
            const uint64_t adesc = (WBK == 64) ? make_kmajor_sw128_desc(w_base + s * Tr::W_STAGE_BYTES)
                                               : make_kmajor_sw64_desc(w_base + s * Tr::W_STAGE_BYTES);
            // token tile: 128-byte rows; this stage covers bytes [sub*WBK*2, +WBK*2) of each row
            const uint64_t bdesc = make_kmajor_sw128_desc(tok_base + kb * Tr::TOK_KB_BYTES) + sub * (WBK * 2 / 16);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < WBK / 16; ++k) {
                // +32 bytes per 16-element k step inside the swizzle row (descriptor units of 16 B => +2)
                umma_bf16<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kh | ws | k) != 0 ? 1u : 0u);
              }
              // W stage reusable once these MMAs retire: the arrival goes to the loading CTAs (cluster ranks 0 and 1)
              // and to this pair's own leader, which arms the next phase (with MC == 1 all of that is this pair)
              umma_commit<CG>(&w_empty[s], MC > 1 ? uint16_t(0x3u | (1u << lcta)) : pair_mask);
              if (tok_last && sub == WPK - 1) umma_commit<CG>(&tok_empty[kb], pair_mask);   // token k-block free
              if (kh == Tr::KS - 1 && ws == Tr::NWSK - 1) umma_commit<CG>(&acc_full[as], pair_mask);
            }
            __syncwarp();
          }
        }
        if constexpr (Tr::KS == 1) ++it;
      }
      if (stall != nullptr && lane == 0) {
        stall[blockIdx.x * 16 + 0] = static_cast<unsigned long long>(clock64() - st_t0);
        stall[blockIdx.x * 16 + 1] = static_cast<unsigned long long>(st_w);
        stall[blockIdx.x * 16 + 2] = static_cast<unsigned long long>(st_tok);
        stall[blockIdx.x * 16 + 3] = static_cast<unsigned long long>(st_acc);
      }
    }
  }

  // ---- teardown -----------------------------------------------------------------------
  __syncwarp();  // single-lane roles (TMA producer, MMA issuer) rejoin their warp before the aligned barriers
  tcgen05_fence_before();
  if constexpr (CL > 1) cluster_sync(); else __syncthreads();
  if (warp == 9) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem_base, Tr::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------

template <class Tr>
static cudaError_t launch_encode_t(const EncoderLaunch& a) {
  auto kern = qc_encode_kernel<Tr>;   // its shared-memory opt-in was set for this device by encoder_prepare_device()
  const int bpc = (a.T + Tr::TB - 1) / Tr::TB;
  constexpr int CL = Tr::CLUSTER;
  int grid = a.num_sms / CL * CL;
  if (a.chunked && a.chunk_row == nullptr) return cudaErrorInvalidValue;
  const long long nblocks = static_cast<long long>(a.R) * bpc;
  if (nblocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int nslots = a.block_map != nullptr ? a.nslots : static_cast<int>(nblocks);
  if (nslots < 0 || nslots > nblocks) return cudaErrorInvalidValue;
  if (nslots == 0) return cudaSuccess;                         // every candidate is empty: K2 writes zero rows
  const long long ncitems = (static_cast<long long>(nslots) + CL - 1) / CL;   // cluster items: CL token blocks each
  if (ncitems * CL < grid) grid = static_cast<int>(ncitems * CL);
  if (grid < CL) grid = CL;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kEncThreads);
  cfg.dynamicSmemBytes = Tr::SMEM_BYTES;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 2 : 1;
  if constexpr (CL > 2) {
    // Clusters of 4 / 8 SMs must sit inside one GPC: the persistent grid is as many clusters as can be co-resident
    // (a wave that does not fit would serialise behind the first).  a.max_clusters caches the occupancy query.
    int fit = a.max_clusters;
    if (fit <= 0) {
      cudaError_t e = cudaOccupancyMaxActiveClusters(&fit, kern, &cfg);
      if (e != cudaSuccess) return e;
    }
    if (fit < 1) return cudaErrorInvalidConfiguration;
    if (grid > fit * CL) cfg.gridDim = dim3(fit * CL);
  }
  const CUtensorMap* tm = Tr::WBK == 64 ? a.w_tmap64 : a.w_tmap32;
  return cudaLaunchKernelEx(&cfg, kern, *tm, *a.e_tmap, a.table, a.token_ids, a.lens, a.part, a.err_flag, a.stall,
                            a.block_map, nslots, a.chunk_row, a.chunked, a.R, a.T, bpc, a.vocab);
}

// The instantiated variants.  {dim, token block, cta group, W stage width, gather mode, pairs per cluster (W multicast)}
#define QC_ENC_VARIANTS(X)                                                                                   \
  X(384, 128, 1, 64, 0, 1)  X(384, 128, 1, 32, 0, 1)  X(384, 128, 1, 64, 1, 1)  X(384, 256, 1, 64, 1, 1)         \
  X(768, 128, 1, 64, 0, 1)  X(768, 128, 1, 32, 0, 1)  X(768, 128, 2, 64, 0, 1)  X(768, 128, 2, 32, 0, 1)         \
  X(768, 112, 2, 64, 0, 1)  X(768, 104, 2, 64, 0, 1)  X(768, 96, 2, 64, 0, 1)                                   \
  X(768, 128, 1, 64, 1, 1)  X(768, 128, 2, 64, 1, 1)  X(768, 128, 2, 32, 1, 1)  X(768, 112, 2, 64, 1, 1)         \
  X(768, 104, 2, 64, 1, 1)  X(768, 96, 2, 64, 1, 1)                                                            \
  X(768, 128, 2, 64, 1, 2)  X(768, 112, 2, 64, 1, 2)  X(768, 104, 2, 64, 1, 2)  X(768, 128, 2, 64, 1, 4)         \
  X(1536, 64, 1, 64, 0, 1)  X(1536, 64, 2, 64, 0, 1)  X(1536, 48, 2, 64, 0, 1)                                  \
  X(1536, 64, 1, 64, 1, 1)  X(1536, 64, 2, 64, 1, 1)  X(1536, 56, 2, 64, 1, 1)  X(1536, 48, 2, 64, 1, 1)         \
  X(1536, 64, 2, 64, 1, 2)  X(1536, 56, 2, 64, 1, 2)  X(1536, 48, 2, 64, 1, 2)                                  \
  X(1536, 64, 2, 64, 1, 4)  X(1536, 56, 2, 64, 1, 4)                                                          \
  X(1536, 128, 2, 64, 1, 1)  X(1536, 112, 2, 64, 1, 1)  X(1536, 96, 2, 64, 1, 1)

#define QC_MATCH(D, TB, CG, WBK, GM, MC) \
  (dim == D && v.token_block == TB && v.cta_group == CG && v.w_bk == WBK && v.gather == GM && (v.mcast < 1 ? 1 : v.mcast) == MC)

// Per-device (= per CUDA context) function state: the > 48 KB dynamic shared-memory opt-in of every tiling of
// `dim`.  qc_ctx_create calls this with the context's device current, so no launch path keeps process-wide
// "already set" flags (a second context on another GPU of the same process would otherwise launch without it).
cudaError_t encoder_prepare_device(int dim) {
#define X(D, TB, CG, WBK, GM, MC)                                                                          \
  if (dim == D) {                                                                                          \
    cudaError_t e = cudaFuncSetAttribute(qc_encode_kernel<EncTraits<D, TB, CG, WBK, GM, MC>>,              \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                                         EncTraits<D, TB, CG, WBK, GM, MC>::SMEM_BYTES);                    \
    if (e != cudaSuccess) return e;                                                                        \
  }
  QC_ENC_VARIANTS(X)
#undef X
  return cudaSuccess;
}

bool encoder_variant_supported(int dim, const EncoderVariant& v) {
#define X(D, TB, CG, WBK, GM, MC) if (QC_MATCH(D, TB, CG, WBK, GM, MC)) return true;
  QC_ENC_VARIANTS(X)
#undef X
  return false;
}

int encoder_variant_w_stages(int dim, const EncoderVariant& v) {
#define X(D, TB, CG, WBK, GM, MC) if (QC_MATCH(D, TB, CG, WBK, GM, MC)) return EncTraits<D, TB, CG, WBK, GM, MC>::W_STAGES;
  QC_ENC_VARIANTS(X)
#undef X
  return 0;
}

EncoderVariant encoder_default_variant(int dim) {
  switch (dim) {
    // fastest measured on B200 (tools/gpu_tune.py, profiles/k1_variants_*_r2.log).  W multicast (mcast 2 / 4) is
    // instantiated and bit-identical but measured SLOWER at every dim (the bound is per-SM ingest, not L2 reads:
    // DESIGN.md section 4), so no default uses it.
    case 384: return EncoderVariant{256, 1, 64, 1, 1};
    case 768: return EncoderVariant{128, 2, 64, 1, 1};
    case 1536: return EncoderVariant{128, 2, 64, 1, 1};      // K-split walk: 1385 vs 1047 TFLOP/s for the best resident tiling
    default: return EncoderVariant{0, 0, 0, 0, 1};
  }
}

// Token-block choice for a given slot count T: the measured throughput of each tiling (profiles/
// k1_variants_*_r1.log, normalised to the fastest) times the fraction of gathered token slots that are
// real (ceil(T/TB)*TB >= T).  T = 512 keeps TB = 128; T = 300 prefers TB = 104 (3 blocks, 4 % padding)
// over TB = 128 (3 blocks, 22 % padding).
EncoderVariant encoder_auto_variant(int dim, int T) {
  struct Cand { EncoderVariant v; float tput; };
  static const Cand c768[] = {{{128, 2, 64, 1, 1}, 1.00f}, {{112, 2, 64, 1, 1}, 0.955f}, {{104, 2, 64, 1, 1}, 0.965f}, {{96, 2, 64, 1, 1}, 0.80f}};
  // profiles/k1_variants_1536_r2.log, run C: the K-split tilings (128 / 112 / 96 tokens per CTA, rows held one K half
  // at a time) against the resident ones (56 / 64 / 48)
  static const Cand c1536[] = {{{128, 2, 64, 1, 1}, 1.00f}, {{112, 2, 64, 1, 1}, 0.92f}, {{96, 2, 64, 1, 1}, 0.85f},
                               {{56, 2, 64, 1, 1}, 0.755f}, {{64, 2, 64, 1, 1}, 0.71f}, {{48, 2, 64, 1, 1}, 0.685f}};
  // dim 384: a 256-token block (UMMA N = 256: half the operand bytes per flop of N = 128) is 5 % faster on full
  // blocks; short candidates keep the 128-token block
  static const Cand c384[] = {{{256, 1, 64, 1, 1}, 1.00f}, {{128, 1, 64, 0, 1}, 0.953f}};
  const Cand* cs = nullptr;
  int n = 0;
  if (dim == 384) { cs = c384; n = 2; }
  else if (dim == 768) { cs = c768; n = 4; }
  else if (dim == 1536) { cs = c1536; n = 6; }
  else return encoder_default_variant(dim);
  EncoderVariant best = cs[0].v;
  float best_score = -1.f;
  for (int i = 0; i < n; ++i) {
    const int tb = cs[i].v.token_block;
    const int padded = ((T + tb - 1) / tb) * tb;
    const float score = cs[i].tput * static_cast<float>(T) / static_cast<float>(padded);
    if (score > best_score) { best_score = score; best = cs[i].v; }
  }
  return best;
}

cudaError_t launch_encode(const EncoderLaunch& a) {
  const int dim = a.dim;
  const EncoderVariant& v = a.variant;
#define X(D, TB, CG, WBK, GM, MC) if (QC_MATCH(D, TB, CG, WBK, GM, MC)) return launch_encode_t<EncTraits<D, TB, CG, WBK, GM, MC>>(a);
  QC_ENC_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace qc
