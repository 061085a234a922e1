// Microbenchmark: how fast can all SMs fill shared memory from an L2-resident matrix with TMA, unicast vs
// cluster multicast?  Decides whether sharing the Z_J / Z^T tile stream between CTA pairs (cluster of 4) can
// lift the ~11 TB/s L2 -> SM rate that bounds the MMD kernels.   nvcc -arch=sm_100a -O3 -o mc tma_multicast.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t par) {
  uint32_t ok;
  asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
  long long t0 = clock64();
  while (!mbar_try(b, par)) if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void arrive_remote(uint32_t a) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int STAGES = 6;
constexpr int TILE_ROWS = 256, TILE_COLS = 32;
constexpr int STAGE_BYTES = TILE_ROWS * TILE_COLS * 4;   // 32 KiB

template <int CSZ, bool MC, int BOXR = 256>
__global__ void __launch_bounds__(64, 1) fill_kernel(const __grid_constant__ CUtensorMap tm, int iters, int nrow_tiles,
                                                      int ncol_tiles, int same_stream) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  const uint32_t rank = (CSZ > 1) ? ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MC ? CSZ : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CSZ > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const int cluster_id = blockIdx.x / CSZ;
  const int base_tile = same_stream ? 0 : cluster_id * 7919;
  if (threadIdx.x == 0) {                      // producer
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[s], ph ^ 1);
      const int t = base_tile + it;
      const int row = (t % nrow_tiles) * TILE_ROWS, col = ((t / nrow_tiles) % ncol_tiles) * TILE_COLS;
      mbar_expect(&full[s], STAGE_BYTES);
      uint8_t *dst = smem + s * STAGE_BYTES;
      if (MC) {
        constexpr int ROWS = TILE_ROWS / CSZ;   // this CTA's slice, multicast to every CTA of the cluster
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(smem_u32(dst + rank * ROWS * TILE_COLS * 4)), "l"((uint64_t)&tm), "r"(smem_u32(&full[s])), "r"(col), "r"(row + (int)rank * ROWS),
                       "h"((uint16_t)((1u << CSZ) - 1)) : "memory");
      } else {
#pragma unroll
        for (int q = 0; q < TILE_ROWS / BOXR; ++q)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(dst + q * BOXR * TILE_COLS * 4)), "l"((uint64_t)&tm), "r"(smem_u32(&full[s])), "r"(col), "r"(row + q * BOXR) : "memory");
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {              // consumer: release the stage as soon as it has landed
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[s], ph);
      if (MC) { for (int r = 0; r < CSZ; ++r) arrive_remote(mapa(smem_u32(&empty[s]), r)); }
      else { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory"); }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (CSZ > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CSZ, bool MC, int BOXR = 256>
void run(const char *name, void *buf, int rows, int cols, PFN_enc enc, int iters, int same) {
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {TILE_COLS, (cuuint32_t)(MC ? TILE_ROWS / CSZ : BOXR)};
  cuuint32_t es[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); exit(1); }
  const int smem = STAGES * STAGE_BYTES + 256;
  auto kern = fill_kernel<CSZ, MC, BOXR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = (148 / CSZ) * CSZ;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int nrt = rows / TILE_ROWS, nct = cols / TILE_COLS;
  for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, kern, tm, iters, nrt, nct, same));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, kern, tm, iters, nrt, nct, same));
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double landed = (double)grid * iters * STAGE_BYTES;
  printf("%-28s cluster=%d %s stream: %7.3f ms  landed %6.2f TB/s  (%5.1f GB/s per SM)\n", name, CSZ, same ? "same " : "own  ", ms,
         landed / ms / 1e9, landed / ms / 1e6 / grid);
}

int main() {
  const int rows = 16384, cols = 512;   // 32 MiB, L2 resident
  void *buf; CK(cudaMalloc(&buf, (size_t)rows * cols * 4)); CK(cudaMemset(buf, 0, (size_t)rows * cols * 4));
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_enc enc = (PFN_enc)fn;
  const int iters = 4096;
  for (int same = 1; same >= 0; --same) {
    run<1, false>("unicast", buf, rows, cols, enc, iters, same);
    run<1, false, 128>("unicast box 128 rows", buf, rows, cols, enc, iters, same);
    run<1, false, 64>("unicast box 64 rows", buf, rows, cols, enc, iters, same);
    run<1, false, 32>("unicast box 32 rows", buf, rows, cols, enc, iters, same);
    run<2, false>("unicast (cluster launch)", buf, rows, cols, enc, iters, same);
    run<2, true>("multicast", buf, rows, cols, enc, iters, same);
    run<4, true>("multicast", buf, rows, cols, enc, iters, same);
    run<8, true>("multicast", buf, rows, cols, enc, iters, same);
  }
  return 0;
}
