// Microbenchmark: can 2-D TMA boxes [32 days x W cells] replay the staging traffic of the
// fused aggregation kernel (many short row runs x 32 day-planes, 4 MB apart) at DRAM speed
// without holding the bytes in registers?  One warp per CTA issues the boxes of a tile
// (512 cells x 32 days = 64 KB) into a ring of `nbuf` shared-memory buffers tracked by
// mbarriers; nobody consumes the data.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bench_micro/tma_box_bw bench_micro/tma_box_bw.cu
//   ./bench_micro/tma_box_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap tm, const int* __restrict__ box_c0,
                                            int nbox, int W, int n_items, int n_spatial, int nbuf,
                                            unsigned long long* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[8];
  const int buf_bytes = nbox * W * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nbuf; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      const int b = k % nbuf;
      if (k >= nbuf) {   // the buffer's previous fill must have landed
        const uint32_t par = ((k / nbuf) - 1) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                       : "=r"(ok) : "r"(s32(&full[b])), "r"(par) : "memory");
      }
      const int sp = item % n_spatial, t0 = (item / n_spatial) * 32;
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[b])), "r"(buf_bytes) : "memory");
      __syncwarp();
      for (int j = lane; j < nbox; j += 32) {
        const int c0 = box_c0[(size_t)sp * nbox + j];
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(s32(smem + (size_t)b * buf_bytes + (size_t)j * W * 128)), "l"(&tm), "r"(c0), "r"(t0),
              "r"(s32(&full[b])) : "memory");
      }
    }
    // drain
    for (int q = 0; q < nbuf && q < k; ++q) {
      const int kk = k - 1 - q, b = kk % nbuf;
      const uint32_t par = (kk / nbuf) & 1;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                     : "=r"(ok) : "r"(s32(&full[b])), "r"(par) : "memory");
    }
    if (lane == 0 && sink && smem[17] == 0xAB) atomicAdd(sink, 1ull);
  }
}

int main(int argc, char** argv) {
  const int T = 1460, NLAT = 720, NLON = 1440;
  const int64_t ncell = (int64_t)NLAT * NLON;
  float* x;
  CK(cudaMalloc(&x, sizeof(float) * ncell * T));
  CK(cudaMemset(x, 0, sizeof(float) * ncell * T));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int n_spatial = 866, n_tb = (T + 31) / 32, n_items = n_spatial * n_tb;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("W cells | boxes/tile | nbuf | CTAs/SM | ms | GB/s staged\n");
  for (int W : {4, 8, 16, 32, 64, 128}) {
    const int nbox = 512 / W;
    // footprint: boxes of one tile sit on consecutive lat rows of a compact patch
    std::vector<int> c0((size_t)n_spatial * nbox);
    srand(1);
    for (int sp = 0; sp < n_spatial; ++sp) {
      const int r0 = rand() % (NLAT - 140), col0 = (rand() % (NLON - 200)) & ~3;
      for (int j = 0; j < nbox; ++j) {
        const int jit = ((rand() % 5) - 2) * 4;
        int c = col0 + 40 + jit;
        c0[(size_t)sp * nbox + j] = (r0 + j) * NLON + c;
      }
    }
    int* d_c0;
    CK(cudaMalloc(&d_c0, c0.size() * 4));
    CK(cudaMemcpy(d_c0, c0.data(), c0.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)ncell, (cuuint64_t)T};
    cuuint64_t strides[1] = {(cuuint64_t)ncell * 4};
    cuuint32_t box[2] = {(cuuint32_t)W, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for W=%d\n", (int)r, W); continue; }
    for (int cfg = 0; cfg < 4; ++cfg) {
      const int nbuf = (cfg == 0) ? 3 : (cfg == 1) ? 2 : (cfg == 2) ? 1 : 1;
      const int ctas = (cfg == 3) ? 3 : (cfg == 2 ? 2 : 1);
      const size_t smem = (size_t)nbuf * 65536;
      CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int it = 0; it < 2; ++it) k_tma<<<sms * ctas, 128, smem>>>(tm, d_c0, nbox, W, n_items, n_spatial, nbuf, sink);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      const int reps = 5;
      for (int it = 0; it < reps; ++it) k_tma<<<sms * ctas, 128, smem>>>(tm, d_c0, nbox, W, n_items, n_spatial, nbuf, sink);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= reps;
      printf("%3d | %3d | %d | %d | %.3f | %.0f\n", W, nbox, nbuf, ctas, ms, (double)n_items * 65536 / ms * 1e-6);
    }
    cudaFree(d_c0);
  }
  return 0;
}
