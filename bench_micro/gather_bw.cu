// Microbenchmark: achievable HBM bandwidth for the staging access pattern of the
// aggregation kernel -- for a spatial tile (ROWS rows x RUN bytes) read the same footprint
// from 32 consecutive day-planes, 16 B per thread, lanes = 8 pieces x 4 days.
// Sweeps loads-in-flight per thread (UNR), warps per SM and run length.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int UNR>
__global__ void gather_kernel(const float* __restrict__ x, long plane_elems, int n_planes, int row_elems,
                              int rows_per_tile, int run_elems, int tiles_x, int tiles_y,
                              float* __restrict__ sink, int land_mod) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int l8 = lane & 7, l4 = lane >> 3;
  const int n_tb = n_planes / 32;
  const long n_tiles = (long)tiles_x * tiles_y * n_tb;
  float acc = 0.f;
  const int pieces_per_row = run_elems / 4;            // 16-byte pieces in one row run
  const int n_pieces = pieces_per_row * rows_per_tile; // per tile per plane
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tb = (int)(tile / ((long)tiles_x * tiles_y));
    const int sp = (int)(tile % ((long)tiles_x * tiles_y));
    if ((sp * 2654435761u >> 8) % 100 >= (unsigned)land_mod) continue;  // ~land fraction
    const int ty = sp / tiles_x, tx = sp % tiles_x;
    // warp w handles days 4*(w%8)+l4, piece subset (w/8) of nwarp/8
    const int day = tb * 32 + 4 * (warp & 7) + l4;
    const float* plane = x + (long)day * plane_elems + (long)ty * rows_per_tile * row_elems + (long)tx * run_elems;
    const int sub = warp >> 3, nsub = nwarp >> 3;
    for (int pg = l8 + 8 * sub; pg < n_pieces; pg += 8 * nsub * UNR) {
      float4 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int q = pg + 8 * nsub * u;
        if (q < n_pieces) {
          const int r = q / pieces_per_row, c = q % pieces_per_row;
          v[u] = ld_stream(plane + (long)r * row_elems + c * 4);
        } else v[u] = make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int nlat = 720, nlon = 1440, n_planes = 736;  // 3.05 GB
  const long plane_elems = (long)nlat * nlon;
  float* x; float* sink;
  cudaMalloc(&x, plane_elems * n_planes * 4);
  cudaMemset(x, 0, plane_elems * n_planes * 4);
  cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("run_B rows unr warps ctas/sm  GB/s  (bytes read)\n");
  for (int run_bytes : {128, 256, 512, 1024}) {
    for (int rows : {27, 8}) {
      const int run_elems = run_bytes / 4;
      const int tiles_x = nlon / run_elems, tiles_y = nlat / rows;
      for (int unr : {4, 8, 16}) {
        for (int warps : {8, 16, 32}) {
          for (int cps : {1, 2}) {
            if (warps * cps > 64) continue;
            const int land = 30;
            // bytes actually read
            long cnt = 0;
            for (int sp = 0; sp < tiles_x * tiles_y; ++sp)
              if ((sp * 2654435761u >> 8) % 100 < (unsigned)land) ++cnt;
            const double bytes = (double)cnt * rows * run_bytes * (n_planes / 32 * 32);
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
              cudaEventRecord(e0);
              const int grid = 148 * cps, block = warps * 32;
#define L(U) gather_kernel<U><<<grid, block>>>(x, plane_elems, n_planes, nlon, rows, run_elems, tiles_x, tiles_y, sink, land)
              if (unr == 4) L(4); else if (unr == 8) L(8); else L(16);
              cudaEventRecord(e1);
              cudaEventSynchronize(e1);
              float t; cudaEventElapsedTime(&t, e0, e1);
              if (rep == 0 || t < ms) ms = t;
            }
            printf("%5d %4d %3d %5d %4d   %7.0f  (%.2f GB)\n", run_bytes, rows, unr, warps, cps,
                   bytes / ms / 1e6, bytes / 1e9);
          }
        }
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
