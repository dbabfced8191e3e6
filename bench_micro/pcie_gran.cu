// At what granularity does a kernel's read of pinned host memory cross PCIe?  Every thread reads 16 bytes at
// i * stride; the time per 16-byte read against the stride tells the transfer unit (ctb_pull_pack reads runs of
// 16-byte pieces whose ends leave sectors / lines partly used).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bench_micro/pcie_gran bench_micro/pcie_gran.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void read16(const uint4* __restrict__ src, size_t stride16, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i * stride16));
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u) *sink = acc;
}

int main() {
  const size_t bytes = 2ull << 30;
  uint4* h; uint4* sink;
  cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
  for (size_t i = 0; i < bytes / 16; ++i) h[i] = make_uint4((unsigned)i, 1, 2, 3);
  cudaMalloc(&sink, 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int stride : {16, 32, 64, 128, 256, 512}) {
    const size_t n = bytes / stride;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      read16<<<148 * 16, 256>>>(h, stride / 16, n, sink);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep == 2)
        printf("stride %4d B: %8.3f ms for %zu reads of 16 B: %6.2f GB/s useful, %6.2f GB/s if 32-B sectors, %6.2f if 64 B, %6.2f if 128 B; span %.2f GB/s\n",
               stride, ms, n, n * 16 / ms / 1e6, n * (double)(stride < 32 ? 16 : 32) / ms / 1e6,
               n * (double)(stride < 64 ? stride : 64) / ms / 1e6, n * (double)(stride < 128 ? stride : 128) / ms / 1e6,
               bytes / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
