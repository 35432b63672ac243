// mma_microbench.cu -- how fast can ONE thread feed tcgen05.mma from shared memory?  (sm_100a)
// Standalone tuning aid (not part of the library): issues ITERS x 4 MMAs (K = 16 each) over operand
// tiles already resident in shared memory, with no loads, no epilogue and no dependencies other than
// a final commit, and reports SM cycles per MMA for cta_group 1 and 2 at several N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I quoracle_b200/csrc tools/mma_microbench.cu -o gpurun_out/mma_microbench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace qc::ptx;

template <int CG, int N>
__global__ void __launch_bounds__(128, 1) mma_loop(int iters, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // A: 8 k-blocks of 128 rows x 128 B (128 KB).  B: 4 k-blocks of (N/CG) rows x 128 B.
  constexpr int A_KB = 8, B_KB = 4;
  constexpr int BROWS = N / CG;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + A_KB * 16384;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (A_KB * 16384 + B_KB * BROWS * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // bf16 pairs, small finite values
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<CG>(&tmem_slot, 512);
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc_bf16_f32(128 * CG, N);
    const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t adesc = make_kmajor_sw128_desc(a0 + (it % A_KB) * 16384);
      const uint64_t bdesc = make_kmajor_sw128_desc(b0 + (it % B_KB) * BROWS * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16<CG>(tmem_base + (it & 1) * 256, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
    }
    const long long t_issue = clock64();
    umma_commit<CG>(&done_bar, 0x1);
    while (!mbar_try_wait(&done_bar, 0)) {
    }
    const long long t1 = clock64();
    out[blockIdx.x * 2 + 0] = static_cast<unsigned long long>(t1 - t0);
    out[blockIdx.x * 2 + 1] = static_cast<unsigned long long>(t_issue - t0);
  }
  __syncwarp();
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

template <int CG, int N>
static void run(int iters, int grid) {
  auto kern = mma_loop<CG, N>;
  const int smem = 1024 + 8 * 16384 + 4 * (N / CG) * 128;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* d_out;
  cudaMalloc(&d_out, sizeof(unsigned long long) * grid * 2);
  cudaMemset(d_out, 0, sizeof(unsigned long long) * grid * 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, iters, d_out);
    cudaEventRecord(e1);
    cudaError_t serr = cudaDeviceSynchronize();
    if (err != cudaSuccess || serr != cudaSuccess) {
      printf("cg=%d N=%d: launch error %s / %s\n", CG, N, cudaGetErrorString(err), cudaGetErrorString(serr));
      return;
    }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(grid * 2);
  cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * grid * 2, cudaMemcpyDeviceToHost);
  double tot = 0, iss = 0;
  int n = 0;
  for (int b = 0; b < grid; ++b)
    if (h[b * 2]) {
      tot += h[b * 2];
      iss += h[b * 2 + 1];
      ++n;
    }
  tot /= n;
  iss /= n;
  const double mmas = 4.0 * iters;
  const double flops = 2.0 * 128 * CG * N * 16 * mmas * n;
  printf("cg=%d M=%d N=%3d grid=%3d: %.1f cycles/MMA (issue-only %.1f), kernel %.3f ms, %.1f TFLOP/s, eff. clock %.0f MHz\n",
         CG, 128 * CG, N, grid, tot / mmas, iss / mmas, ms, flops / (ms * 1e-3) / 1e12, tot / (ms * 1e-3) / 1e6);
  cudaFree(d_out);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs, iters=%d (x4 MMAs, K=16 each)\n", p.name, sms, iters);
  run<1, 64>(iters, sms);
  run<1, 128>(iters, sms);
  run<1, 256>(iters, sms);
  run<2, 128>(iters, sms & ~1);
  run<2, 192>(iters, sms & ~1);
  run<2, 256>(iters, sms & ~1);
  run<1, 128>(iters, 1);
  run<2, 256>(iters, 2);
  return 0;
}
