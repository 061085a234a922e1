// Microbenchmark: issue rate of tcgen05.mma (SS mode, 128-byte swizzled K-major operands in shared memory) by
// instruction shape, kind and cta_group.  Answers: does an M=128 cta_group::2 MMA (64 rows per CTA) run at the
// same MAC rate as an M=256 one (128 rows per CTA)?  How much faster is kind::f16 than kind::tf32 when both
// operands come from shared memory?   Operand data is zero (the rate does not depend on it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -o mma_rate mma_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"
using namespace edrl::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NBUF = 4;

__device__ __forceinline__ void mma_f16_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// CG: cta_group; M, N: instruction shape (over the pair when CG == 2); F16: kind::f16 instead of kind::tf32;
// BSHARE: every MMA reads the same B buffer (isolates the A-side cost)
template <int CG, int M, int N, bool F16>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long *cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_ROWS = M / CG, B_ROWS = N / CG;
  constexpr int A_BYTES = A_ROWS * 128, B_BYTES = B_ROWS * 128;
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + NBUF * A_BYTES;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < NBUF * (A_BYTES + B_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_pair(&tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    constexpr uint32_t idesc = F16 ? make_idesc_f16(M, N) : make_idesc_tf32(M, N);
    const uint32_t a_addr = smem_u32(a_base), b_addr = smem_u32(b_base);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int buf = it % NBUF;
      const uint64_t a_d = make_kmajor_sw128_desc(a_addr + buf * A_BYTES);
      const uint64_t b_d = make_kmajor_sw128_desc(b_addr + buf * B_BYTES);
      const uint32_t d = tmem + ((N <= 256) ? (uint32_t)((it >> 4) & 1) * 256u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(k * 2);
        if (CG == 2) {
          if (F16) mma_f16_ss_pair_elect(d, a_d + adv, b_d + adv, idesc, 1u);
          else mma_tf32_ss_pair_elect(d, a_d + adv, b_d + adv, idesc, 1u);
        } else {
          if (F16) mma_f16_ss_elect(d, a_d + adv, b_d + adv, idesc, 1u);
          else mma_tf32_ss_elect(d, a_d + adv, b_d + adv, idesc, 1u);
        }
      }
    }
    if (CG == 2) mma_commit_pair_elect(&done_bar);
    else mma_commit_elect(&done_bar);
  }
  if (CG == 2) {
    // the multicast commit arrives on both CTAs' barriers
    if (threadIdx.x == 0) mbar_wait(&done_bar, 0);
  } else if (threadIdx.x == 0) {
    mbar_wait(&done_bar, 0);
  }
  if (warp == 0 && rank == 0) {
    mbar_wait(&done_bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && cycles_out) cycles_out[blockIdx.x / CG] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

template <int CG, int M, int N, bool F16>
static void run(const char *name, int iters, int nsm) {
  constexpr int A_ROWS = M / CG, B_ROWS = N / CG;
  const int smem_bytes = NBUF * (A_ROWS + B_ROWS) * 128 + 1024;
  auto kern = mma_rate_kernel<CG, M, N, F16>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  long long *cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nsm / CG * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) CK(cudaLaunchKernelEx(&cfg, kern, iters, cyc));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, kern, iters, cyc));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[148];
  CK(cudaMemcpy(h, cyc, sizeof(long long) * (nsm / CG), cudaMemcpyDeviceToHost));
  double csum = 0;
  for (int i = 0; i < nsm / CG; ++i) csum += (double)h[i];
  const double cyc_avg = csum / (nsm / CG);
  const int kper = F16 ? 16 : 8;
  const double mmas = (double)iters * 4;
  const double flops = 2.0 * M * N * kper * mmas * (nsm / CG);
  const double smem_bytes_per_mma_cta = (double)(A_ROWS + B_ROWS) * 32;
  printf("%-34s  %8.1f clk/MMA  %7.1f MAC/clk/SM  %6.1f B/clk/SM smem  %8.1f TFLOP/s (event, %.3f ms)\n", name,
         cyc_avg / mmas, (double)M * N * kper / CG / (cyc_avg / mmas), smem_bytes_per_mma_cta / (cyc_avg / mmas),
         flops / (ms * 1e-3) / 1e12, ms);
  CK(cudaFree(cyc));
}

int main() {
  int dev = 0, nsm = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs: %d\n", nsm);
  const int iters = 20000;
  run<2, 128, 128, false>("tf32 cg2 M=128 N=128 (64 rows/CTA)", iters, nsm);
  run<2, 128, 256, false>("tf32 cg2 M=128 N=256 (64 rows/CTA)", iters, nsm);
  run<2, 256, 128, false>("tf32 cg2 M=256 N=128", iters, nsm);
  run<2, 256, 256, false>("tf32 cg2 M=256 N=256", iters, nsm);
  run<1, 64, 256, false>("tf32 cg1 M=64  N=256", iters, nsm);
  run<1, 128, 128, false>("tf32 cg1 M=128 N=128", iters, nsm);
  run<1, 128, 256, false>("tf32 cg1 M=128 N=256", iters, nsm);
  run<2, 128, 128, true>("f16  cg2 M=128 N=128 (64 rows/CTA)", iters, nsm);
  run<2, 128, 256, true>("f16  cg2 M=128 N=256 (64 rows/CTA)", iters, nsm);
  run<2, 256, 128, true>("f16  cg2 M=256 N=128", iters, nsm);
  run<2, 256, 256, true>("f16  cg2 M=256 N=256", iters, nsm);
  run<1, 128, 256, true>("f16  cg1 M=128 N=256", iters, nsm);
  return 0;
}
