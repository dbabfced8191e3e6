// Dispatch of ctb_aggregate / ctb_aggregate_grouped, and the kernels that are not the
// streaming kernel (ctb_stream.cu):
//
//   agg_snyder_kernel  TIME_MAJOR (tasmin, tasmax) -> Snyder EDD / GDD fused into the gather
//                      (transformations.py:69-89, 139-141 + aggregations.py:27, 75-82).  fp64
//                      ALU bound: CTAs of 8 warps with up to 128 registers, two per SM; the
//                      footprint is staged with 16-byte loads through registers into a
//                      TRANSPOSED tile sx[input][cell][day] (row stride 33 => conflict-free);
//                      one warp per region, lane = day, regions taken from a shared counter.
//   agg_direct_kernel  CELL_MAJOR input [lat][lon][T] (the reference test fixture), and the
//                      fallback for planes that are not 16-byte aligned: one warp per
//                      (region, 32-day tile), lane = day.
//   agg_fixup_kernel   regions split over several bundles, and regions without kept rows:
//                      out[r][t] = (sum of the region's partial rows, fixed order) / den[r].
//   agg_group_finish_kernel  fused time reduction: per-tile partial sums -> output columns.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ctb_internal.cuh"

namespace {

template <int KIND>
struct NIn { static constexpr int v = (KIND == CTB_TR_EDD || KIND == CTB_TR_GDD) ? 2 : 1; };

// ---- mbarrier + bulk async copy (TMA 1-D; SASS: UBLKCP / SYNCS) ---------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n"
      "@p bra DONE_%=;\n"
      "nanosleep.u32 128;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
template <typename T>
__device__ __forceinline__ T lds_val(uint32_t addr) {
  T v;
  if constexpr (sizeof(T) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// one CSR entry: acc_j += w * f_j(x), NaN products skipped (skipna sum, aggregations.py:78)
template <typename TIN, int KIND, int NOUT>
__device__ __forceinline__ void accumulate(const CtbTr& tr, double w, TIN r0, TIN r1, double (&acc)[NOUT]) {
  double f[NOUT];
  ctb_apply<KIND, NOUT>(tr, (double)r0, (double)r1, f);
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const double p = w * f[j];
    if (p == p) acc[j] += p;
  }
}

// time group of the lane's day (fused time reduction), and of the tile's first day
__device__ __forceinline__ void lane_group(const AggArgs& a, int t, bool valid, int& tg, int& tg0) {
  tg = -1;
  tg0 = 0;
  if (a.tgroup) {
    tg = valid ? __ldg(a.tgroup + a.t_off + t) : -1;
    tg0 = __shfl_sync(0xffffffffu, tg, 0);
  }
}

template <typename TIN, int KIND, int NOUT, bool VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) agg_snyder_kernel(const AggArgs a) {
  constexpr int NIN = NIn<KIND>::v;
  constexpr int TILE_LOADS = 8;                 // 16-byte loads per thread in flight
  constexpr int S = CTB_S;
  constexpr int HALVES = sizeof(TIN) / 4;       // 16-byte units per piece-day of one input
  constexpr int CPU = 16 / sizeof(TIN);         // cells per unit
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_unit;
  __shared__ int s_seg_next;   // next region of the tile nobody has taken yet

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* const s_blob = smem_raw + a.tile_stride;
  if (tid == 0) mbar_init(&s_bar, 1);

  // Work unit = (bundle, chunk of `chunk_tb` consecutive 32-day blocks), handed out by an
  // atomic counter in chunk-major order; the bundle's metadata is fetched once per unit.
  for (int n_done = 0;; ++n_done) {
    __syncthreads();   // previous unit fully reduced: tile, blob and s_unit may be reused
    if (tid == 0) {
      s_unit = atomicAdd(a.work_counter, 1);
      if (s_unit < a.n_items) {
        const int4 d = __ldg(a.b_desc + s_unit % a.n_bundles);
        const int64_t o = ((int64_t)(uint32_t)d.y << 32) | (uint32_t)d.x;
        bulk_g2s(s_blob, a.blob + o, (uint32_t)d.z, &s_bar);
      }
    }
    __syncthreads();
    const int unit = s_unit;
    if (unit >= a.n_items) break;
    const int tb_begin = (unit / a.n_bundles) * a.chunk_tb;
    const int tb_end = min(tb_begin + a.chunk_tb, a.n_tb);
    mbar_wait(&s_bar, n_done & 1);
    const CtbBlobHeader H = *reinterpret_cast<const CtbBlobHeader*>(s_blob);
    const int* s_piece = reinterpret_cast<const int*>(s_blob + sizeof(CtbBlobHeader));
    const unsigned char* mb = s_blob + H.bytes_a;
    const int nP = H.n_pieces;

    for (int tb = tb_begin; tb < tb_end; ++tb) {
      const int t0 = tb * CTB_TB;
      if (tb != tb_begin) __syncthreads();   // tile buffer free again
      if (tid == 0) s_seg_next = 0;
      // ---------------- stage: [day][piece] global  ->  [input][cell][day] shared -------------
      {
        const int l8 = lane & 7, l4 = lane >> 3;
        const int dl = (warp & 7) * 4 + l4;            // day within the tile
        constexpr int NSUB = (THREADS / 32) / 8;       // warps sharing one 4-day group
        const int sub = warp >> 3;
        const int t = t0 + dl;
        const int nPH = nP * HALVES;
        const int n_units = nPH * NIN;                 // unit index: [input][piece][half]
        if (t < a.T) {
          const int64_t tp = a.tix ? a.tix[t] : t;
          const TIN* p0 = reinterpret_cast<const TIN*>(a.x0) + tp * a.stride;
          const TIN* p1 = NIN == 2 ? reinterpret_cast<const TIN*>(a.x1) + tp * a.stride : p0;
          asm volatile("" : "+l"(p0));   // keep the day's base pointers in registers
          if constexpr (NIN == 2) asm volatile("" : "+l"(p1));
          TIN* sx = reinterpret_cast<TIN*>(smem_raw) + dl;
          for (int g0 = l8 + 8 * sub; g0 < n_units; g0 += 8 * NSUB * TILE_LOADS) {
            int off[TILE_LOADS];
            uint32_t v[TILE_LOADS][4];
#pragma unroll
            for (int u = 0; u < TILE_LOADS; ++u) {      // all index reads first, then all loads
              const int g = g0 + 8 * NSUB * u;
              if (g < n_units) {
                const int in = (NIN == 2 && g >= nPH) ? 1 : 0, r = g - (in ? nPH : 0);
                off[u] = s_piece[r / HALVES] * CTB_PIECE + (r % HALVES) * CPU;
              }
            }
#pragma unroll
            for (int u = 0; u < TILE_LOADS; ++u) {
              const int g = g0 + 8 * NSUB * u;
              if (g < n_units) {
                const TIN* src = (NIN == 2 && g >= nPH) ? p1 + off[u] : p0 + off[u];
                if constexpr (VEC) {
                  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                               : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]) : "l"(src));
                } else {
                  TIN tv[CPU];
#pragma unroll
                  for (int q = 0; q < CPU; ++q) tv[q] = (off[u] + q < a.ncell) ? __ldg(src + q) : TIN(0);
                  if constexpr (sizeof(TIN) == 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[u][q] = __float_as_uint((float)tv[q]);
                  } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                      const long long bb = __double_as_longlong((double)tv[q]);
                      v[u][2 * q] = (uint32_t)bb; v[u][2 * q + 1] = (uint32_t)(bb >> 32);
                    }
                  }
                }
              }
            }
#pragma unroll
            for (int u = 0; u < TILE_LOADS; ++u) {
              const int g = g0 + 8 * NSUB * u;
              if (g < n_units) {
                // unit g covers cells [g*CPU, g*CPU + CPU) of the [input][cell] row space
                TIN* sd = sx + (size_t)g * CPU * S;
                if constexpr (sizeof(TIN) == 4) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) sd[q * S] = __uint_as_float(v[u][q]);
                } else {
#pragma unroll
                  for (int q = 0; q < 2; ++q)
                    sd[q * S] = __longlong_as_double(((long long)v[u][2 * q + 1] << 32) | v[u][2 * q]);
                }
              }
            }
          }
        }
      }
      __syncthreads();

      // ---------------- gather + segmented weighted sum: warp = region, lane = day ------------
      const CtbSeg* segs = reinterpret_cast<const CtbSeg*>(mb + H.off_seg);
      const CtbEnt* ENT = reinterpret_cast<const CtbEnt*>(mb + H.off_ent);
      const uint32_t sb0 = smem_u32(smem_raw) + lane * (uint32_t)sizeof(TIN);
      const uint32_t sb1 = sb0 + (uint32_t)(nP * CTB_PIECE * S) * (uint32_t)sizeof(TIN);
      const int t = t0 + lane;
      const bool valid = t < a.T;
      int tg, tg0;
      lane_group(a, t, valid, tg, tg0);
      // the planner's column offsets are row-major (column * elem_bytes): x 33 for this tile
      auto at0 = [&](uint32_t o) { return lds_val<TIN>(sb0 + o * (uint32_t)S); };
      auto at1 = [&](uint32_t o) { return lds_val<TIN>(sb1 + o * (uint32_t)S); };
      for (int s = warp; s < H.n_seg;) {
        const CtbSeg sg = segs[s];
        double acc[NOUT], acc2[NOUT];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) acc[j] = acc2[j] = 0.0;
        // Quads of entries, software-pipelined: the metadata and the staged values of quad c+1
        // are loaded before quad c is accumulated.  The padding of the last quad has weight 0
        // and is masked out (0 * NaN must not reach the sum).
        const int e0 = (int)sg.e0_4 * 4;
        const int n_chunks = ((int)sg.n + 3) >> 2;
        const int last_valid = (int)sg.n - 4 * (n_chunks - 1);   // 1..4 entries in the last quad
        double wA[4];
        TIN xA[4], yA[4];
        auto fetch = [&](int e, double (&w)[4], TIN (&x0)[4], TIN (&x1)[4]) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 m = *reinterpret_cast<const uint4*>(ENT + e + q);
            w[q] = __hiloint2double((int)m.y, (int)m.x);
            x0[q] = at0(m.z);
            x1[q] = at1(m.z);
          }
        };
        if (n_chunks > 0) fetch(e0, wA, xA, yA);
        for (int c = 0; c < n_chunks; ++c) {
          double wB[4];
          TIN xB[4], yB[4];
          const bool more = c + 1 < n_chunks;
          if (more) fetch(e0 + 4 * (c + 1), wB, xB, yB);
          if (more || last_valid == 4) {
            accumulate<TIN, KIND, NOUT>(a.tr, wA[0], xA[0], yA[0], acc);
            accumulate<TIN, KIND, NOUT>(a.tr, wA[1], xA[1], yA[1], acc2);
            accumulate<TIN, KIND, NOUT>(a.tr, wA[2], xA[2], yA[2], acc);
            accumulate<TIN, KIND, NOUT>(a.tr, wA[3], xA[3], yA[3], acc2);
          } else {
            accumulate<TIN, KIND, NOUT>(a.tr, wA[0], xA[0], yA[0], acc);
            if (last_valid > 1) accumulate<TIN, KIND, NOUT>(a.tr, wA[1], xA[1], yA[1], acc2);
            if (last_valid > 2) accumulate<TIN, KIND, NOUT>(a.tr, wA[2], xA[2], yA[2], acc);
          }
          if (more) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { wA[q] = wB[q]; xA[q] = xB[q]; yA[q] = yB[q]; }
          }
        }
        double v[NOUT];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) v[j] = acc[j] + acc2[j];
        ctb_emit<NOUT>(a, sg.target, sg.rden, v, lane, t, valid, tb, tg, tg0);
        int nx = 0;
        if (lane == 0) nx = (THREADS / 32) + atomicAdd(&s_seg_next, 1);
        s = __shfl_sync(0xffffffffu, nx, 0);
      }
    }   // tiles of the unit
  }   // units
}

// Regions split over several bundles (and regions with no kept rows):
// out[r][t] = (sum of the region's partial rows, fixed order) / den[r]; with a time reduction
// out[r][g] = sum over the days of column g of that.
__global__ void agg_fixup_kernel(const AggArgs a, int n_out) {
  const int i = blockIdx.x;
  const int r = a.split_region[i];
  const int s0 = a.split_slot_ptr[i], s1 = a.split_slot_ptr[i + 1];
  const double d = a.den[r];
  for (int j = 0; j < n_out; ++j) {
    if (!a.tgroup) {
      for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < a.T; t += gridDim.y * blockDim.x) {
        double s = 0.0;
        for (int q = s0; q < s1; ++q) s += a.scratch[((size_t)j * a.n_scratch + q) * a.scratch_ld + t];
        a.out[((size_t)j * a.R + r) * a.out_ld + t] = s / d;
      }
    } else {
      for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < a.n_groups; g += gridDim.y * blockDim.x) {
        double sum = 0.0;
        for (int t = a.g_t_lo[g]; t < a.g_t_hi[g]; ++t) {
          double s = 0.0;
          for (int q = s0; q < s1; ++q) s += a.scratch[((size_t)j * a.n_scratch + q) * a.scratch_ld + t];
          sum += s / d;
        }
        a.out[((size_t)j * a.R + r) * a.out_ld + g] = sum;
      }
    }
  }
}

// fused time reduction: out[j][r][g] = sum over the tiles that touch column g of their partial
// sums, in tile order
__global__ void agg_group_finish_kernel(const AggArgs a, int n_out) {
  const int64_t n = (int64_t)n_out * a.R * a.n_groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % a.n_groups);
    const int64_t jr = i / a.n_groups;
    const int t_lo = a.g_t_lo[g], t_hi = a.g_t_hi[g];
    double s = 0.0;
    if (t_hi > t_lo) {
      for (int tb = t_lo / CTB_TB; tb <= (t_hi - 1) / CTB_TB; ++tb) {
        const int k = g - a.tgroup[tb * CTB_TB];
        s += a.gpart[((size_t)jr * a.g_ntb + tb) * a.gk + k];
      }
    }
    a.out[(size_t)jr * a.out_ld + g] = s;
  }
}

// ---- direct kernel: warp per (region, 32-day tile) ---------------------------
template <typename TIN, int KIND, int NOUT, int LAYOUT>
__global__ void __launch_bounds__(256) agg_direct_kernel(const AggArgs a) {
  constexpr int NIN = NIn<KIND>::v;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (a.T + 31) / 32;
  const int64_t n_work = (int64_t)a.R * n_tiles;
  const TIN* __restrict__ X0 = reinterpret_cast<const TIN*>(a.x0);
  const TIN* __restrict__ X1 = reinterpret_cast<const TIN*>(a.x1);
  for (int64_t wk = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); wk < n_work;
       wk += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int r = (int)(wk / n_tiles);
    const int tb = (int)(wk % n_tiles);
    const int t = tb * 32 + lane;
    const bool tv = t < a.T;
    const int64_t tp = tv ? (a.tix ? a.tix[t] : t) : 0;
    int tg, tg0;
    lane_group(a, t, tv, tg, tg0);
    double acc[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) acc[j] = 0.0;
    const int e0 = a.row_ptr[r], e1 = a.row_ptr[r + 1];
    for (int e = e0; e < e1; ++e) {
      const double w = __ldg(a.w + e);
      const int64_t c = __ldg(a.col + e);
      const int64_t off = (LAYOUT == CTB_LAYOUT_CELL_MAJOR) ? c * a.stride + tp : tp * a.stride + c;
      double x0 = 0.0, x1 = 0.0;
      if (tv) {
        x0 = (double)__ldg(X0 + off);
        if constexpr (NIN == 2) x1 = (double)__ldg(X1 + off);
      }
      accumulate<double, KIND, NOUT>(a.tr, w, x0, x1, acc);
    }
    // 1/den: the same normalisation as the staged kernels (0 * inf = NaN, x * inf = +-inf)
    ctb_emit<NOUT>(a, r, 1.0 / a.den[r], acc, lane, t, tv, tb, tg, tg0);
  }
}

// ------------------------------------------------------------- dispatch -----
template <typename TIN, int KIND, int NOUT>
int launch_snyder(const ctb_plan* P, AggArgs a, bool vec, cudaStream_t st) {
  constexpr int NIN = NIn<KIND>::v;
  constexpr int THREADS = 256;
  static int n_sm[64] = {0};
  static bool attr_set[64][2] = {{false}};
  const int dev = P->device & 63;
  const size_t tile = ((size_t)NIN * P->info.max_bundle_cells * CTB_S * sizeof(TIN) + 127) & ~(size_t)127;
  const size_t smem = tile + CTB_META_CAP;
  const size_t smem_cap = 164 * 1024 / CTB_CTAS_PER_SM - 1024 - 512;
  if (smem > smem_cap) {
    ctb_set_error("plan's bundles need %zu bytes of staging (cap %zu): build it with stage_bytes = n_in * elem_bytes",
                  smem, smem_cap);
    return CTB_ERR_UNSUPPORTED;
  }
  if (!n_sm[dev]) CTB_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, P->device));
  const int n_tb = (a.T + CTB_TB - 1) / CTB_TB;
  int chunk_tb = 4;
  const int n_chunks = (n_tb + chunk_tb - 1) / chunk_tb;
  chunk_tb = n_chunks ? (n_tb + n_chunks - 1) / n_chunks : 1;
  const int64_t n_units = (int64_t)P->n_bundles * n_chunks;
  if (n_units >= (1ll << 31)) { ctb_set_error("too many work units"); return CTB_ERR_UNSUPPORTED; }
  a.n_bundles = P->n_bundles; a.n_items = (int)n_units; a.n_tb = n_tb; a.chunk_tb = std::max(chunk_tb, 1);
  a.tile_stride = (int)tile;
  a.work_counter = P->d_work_counter + P->work_counter_slot.fetch_add(1) % CTB_N_WORK_COUNTERS;
  if (n_units > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(n_units, (int64_t)n_sm[dev] * CTB_CTAS_PER_SM);
    CTB_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int), st));
    auto go = [&](auto k) -> int {
      if (!attr_set[dev][vec ? 1 : 0]) {
        // 164 KB of shared memory per SM: the L1 that is left holds the loads in flight
        CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 72));
        attr_set[dev][vec ? 1 : 0] = true;
      }
      k<<<grid, THREADS, smem, st>>>(a);
      return CTB_OK;
    };
    const int rc = vec ? go(agg_snyder_kernel<TIN, KIND, NOUT, true, THREADS>)
                       : go(agg_snyder_kernel<TIN, KIND, NOUT, false, THREADS>);
    if (rc) return rc;
    CTB_LAUNCH_CHECK();
  }
  return CTB_OK;
}

template <typename TIN, int KIND, int NOUT>
int launch_direct(const AggArgs& a, int layout, cudaStream_t st) {
  AggArgs b = a;
  b.n_tb = (a.T + CTB_TB - 1) / CTB_TB;
  const int64_t n_work = (int64_t)a.R * b.n_tb;
  if (n_work == 0) return CTB_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((n_work + 7) / 8, 148 * 64);
  if (layout == CTB_LAYOUT_CELL_MAJOR)
    agg_direct_kernel<TIN, KIND, NOUT, CTB_LAYOUT_CELL_MAJOR><<<grid, 256, 0, st>>>(b);
  else
    agg_direct_kernel<TIN, KIND, NOUT, CTB_LAYOUT_TIME_MAJOR><<<grid, 256, 0, st>>>(b);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

// variant 1 = staged (streaming kernel for IDENTITY / POLY, Snyder kernel for EDD / GDD), 2 = direct
template <typename TIN, int KIND, int NOUT>
int run(const ctb_plan* P, const AggArgs& a, int layout, int variant, bool vec, cudaStream_t st) {
  if (variant == 2) return launch_direct<TIN, KIND, NOUT>(a, layout, st);
  if constexpr (KIND == CTB_TR_EDD || KIND == CTB_TR_GDD) return launch_snyder<TIN, KIND, NOUT>(P, a, vec, st);
  else return ctb_launch_stream(P, a, std::is_same<TIN, float>::value ? CTB_F32 : CTB_F64, KIND, NOUT, st);
}

template <typename TIN, int KIND>
int run_nout(const ctb_plan* P, const AggArgs& a, int layout, int variant, bool vec, int n_out,
             cudaStream_t st) {
  switch (n_out) {
    case 1: return run<TIN, KIND, 1>(P, a, layout, variant, vec, st);
    case 2: return run<TIN, KIND, 2>(P, a, layout, variant, vec, st);
    case 3: return run<TIN, KIND, 3>(P, a, layout, variant, vec, st);
    case 4: return run<TIN, KIND, 4>(P, a, layout, variant, vec, st);
  }
  ctb_set_error("n_out=%d unsupported", n_out);
  return CTB_ERR_INVALID;
}

template <typename TIN>
int run_kind(const ctb_plan* P, const AggArgs& a, int layout, int variant, bool vec, int kind,
             int n_out, cudaStream_t st) {
  switch (kind) {
    case CTB_TR_IDENTITY: return run<TIN, CTB_TR_IDENTITY, 1>(P, a, layout, variant, vec, st);
    case CTB_TR_POLY: return run_nout<TIN, CTB_TR_POLY>(P, a, layout, variant, vec, n_out, st);
    case CTB_TR_EDD: return run_nout<TIN, CTB_TR_EDD>(P, a, layout, variant, vec, n_out, st);
    case CTB_TR_GDD: return run_nout<TIN, CTB_TR_GDD>(P, a, layout, variant, vec, n_out, st);
  }
  ctb_set_error("transform=%d unsupported", kind);
  return CTB_ERR_INVALID;
}

size_t scratch_bytes(const ctb_plan* plan, int64_t T, int n_out) {
  return (size_t)plan->n_scratch * (size_t)T * (size_t)n_out * sizeof(double);
}
size_t gpart_bytes(const ctb_plan* plan, const ctb_time_groups* g, int n_out) {
  const size_t n_tb = (size_t)((g->T + CTB_TB - 1) / CTB_TB);
  return (size_t)n_out * (size_t)plan->R * n_tb * (size_t)g->gk * sizeof(double);
}

int aggregate_impl(const ctb_plan* P, const void* x0, const void* x1, int dtype, int layout,
                   int64_t stride, const int32_t* time_index, int64_t T, int transform,
                   const double* params, int n_params, int n_out, const ctb_time_groups* G,
                   int64_t t_begin, int flush, double* out,
                   int64_t out_ld, void* workspace, size_t workspace_bytes, int variant, void* stream) {
  const char* fn = G ? "ctb_aggregate_grouped" : "ctb_aggregate";
  if (!P || !x0 || (!out && T > 0 && P->R > 0)) { ctb_set_error("%s: null argument", fn); return CTB_ERR_INVALID; }
  const int64_t n_cols = G ? G->n_groups : T;
  if (out_ld == 0) out_ld = n_cols;
  if (T < 0 || T >= (1ll << 31) || stride < 0 || out_ld < n_cols) { ctb_set_error("%s: bad T/stride/out_ld", fn); return CTB_ERR_INVALID; }
  if (G && (t_begin < 0 || t_begin % CTB_TB != 0 || t_begin + T > G->T || G->device != P->device)) {
    ctb_set_error("%s: window [%lld, %lld) does not fit time groups of %lld days on device %d (t_begin must be a multiple of %d)",
                  fn, (long long)t_begin, (long long)(t_begin + T), (long long)G->T, G->device, CTB_TB);
    return CTB_ERR_INVALID;
  }
  if (dtype != CTB_F32 && dtype != CTB_F64) { ctb_set_error("dtype=%d unsupported", dtype); return CTB_ERR_INVALID; }
  if (layout != CTB_LAYOUT_TIME_MAJOR && layout != CTB_LAYOUT_CELL_MAJOR) { ctb_set_error("layout=%d unsupported", layout); return CTB_ERR_INVALID; }
  AggArgs a{};
  int rc = ctb_pack_transform(transform, params, n_params, n_out, &a.tr);
  if (rc) return rc;
  if (ctb_tr_nin(transform) == 2 && !x1) { ctb_set_error("transform needs two inputs"); return CTB_ERR_INVALID; }
  variant &= 0xff;   // bit 8 (inputs in mapped host memory) needs no special handling
  if (variant == 0) variant = (layout == CTB_LAYOUT_TIME_MAJOR) ? 1 : 2;
  if (variant != 2 && layout != CTB_LAYOUT_TIME_MAJOR) { ctb_set_error("staged variant needs TIME_MAJOR input"); return CTB_ERR_INVALID; }
  if (variant != 1 && variant != 2) { ctb_set_error("variant=%d unsupported", variant); return CTB_ERR_INVALID; }
  const size_t es = dtype == CTB_F32 ? 4 : 8;
  const bool vec = (P->ncell % CTB_PIECE == 0) && ((stride * es) % 16 == 0) &&
                   ((uintptr_t)x0 % 16 == 0) && (!x1 || (uintptr_t)x1 % 16 == 0);
  const bool snyder = transform == CTB_TR_EDD || transform == CTB_TR_GDD;
  // the streaming kernel copies 16-byte units: planes that are not 16-byte aligned take the direct kernel
  if (variant == 1 && !snyder && !vec) variant = 2;
  if (variant == 1 && P->elem_bytes != (int)es) {
    ctb_set_error("plan was built for %d-byte elements, input has %d-byte elements: rebuild it with "
                  "elem_bytes=%d", P->elem_bytes, (int)es, (int)es);
    return CTB_ERR_UNSUPPORTED;
  }
  if (variant == 1 && P->stage_bytes < ctb_tr_nin(transform) * (int)es) {
    ctb_set_error("plan stages %d bytes per gridcell-day, the transform needs %d: rebuild it with "
                  "stage_bytes_per_cell_day=%d", P->stage_bytes, ctb_tr_nin(transform) * (int)es,
                  ctb_tr_nin(transform) * (int)es);
    return CTB_ERR_UNSUPPORTED;
  }
  // workspace: [partial rows of split regions][per-tile partial sums of the time reduction]
  const size_t need_s = variant == 1 ? scratch_bytes(P, G ? G->T : T, n_out) : 0;
  const size_t need_g = G ? gpart_bytes(P, G, n_out) : 0;
  if (need_s + need_g > 0 && (!workspace || workspace_bytes < need_s + need_g)) {
    ctb_set_error("workspace of %zu bytes required, got %zu", need_s + need_g, workspace ? workspace_bytes : (size_t)0);
    return CTB_ERR_INVALID;
  }
  if (P->R == 0 || (T == 0 && !(G && flush))) return CTB_OK;

  CtbDeviceGuard guard(P->device);
  if (guard.err != cudaSuccess) { ctb_set_error("cudaSetDevice(%d) failed: %s", P->device, cudaGetErrorString(guard.err)); return CTB_ERR_CUDA; }
  a.x0 = x0; a.x1 = x1; a.stride = stride; a.tix = time_index; a.T = (int)T; a.out_ld = out_ld; a.ncell = P->ncell;
  a.R = P->R; a.out = out; a.scratch = (double*)workspace; a.n_scratch = P->n_scratch;
  a.den = P->d_den;
  a.blob = P->d_blob; a.b_desc = P->d_b_desc; a.unit_tab = P->d_unit_tab; a.row_ptr = P->d_row_ptr; a.col = P->d_col;
  a.w = P->d_w; a.split_region = P->d_split_region; a.split_slot_ptr = P->d_split_slot_ptr;
  a.n_split = P->n_split;
  a.n_tb = (int)((T + CTB_TB - 1) / CTB_TB);
  a.scratch_ld = T;
  if (G) {
    a.tgroup = G->d_group; a.gk = G->gk; a.n_groups = G->n_groups; a.g_t_lo = G->d_t_lo; a.g_t_hi = G->d_t_hi;
    a.g_ntb = (int)((G->T + CTB_TB - 1) / CTB_TB); a.t_off = (int)t_begin; a.scratch_ld = G->T;
    a.gpart = reinterpret_cast<double*>(static_cast<char*>(workspace) + need_s);
  }
#ifdef CTB_EXPERIMENT
  if (const char* e = getenv("CTB_KNOBS")) a.knobs = atoi(e);
  if (const char* e = getenv("CTB_STAGES")) a.n_stages = atoi(e);
  if (const char* e = getenv("CTB_CHUNK_TB")) a.chunk_tb = atoi(e);
#endif
  cudaStream_t st = (cudaStream_t)stream;
  if (T > 0) {
    rc = dtype == CTB_F32 ? run_kind<float>(P, a, layout, variant, vec, transform, n_out, st)
                          : run_kind<double>(P, a, layout, variant, vec, transform, n_out, st);
    if (rc) return rc;
  }
  if (G && !flush) return CTB_OK;
  if (G) {
    a.T = (int)G->T;
    const int64_t n = (int64_t)n_out * P->R * G->n_groups;
    agg_group_finish_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(a, n_out);
    CTB_LAUNCH_CHECK();
  }
  if (variant == 1 && P->n_split > 0) {
    const int64_t cols = G ? G->n_groups : T;
    const dim3 g2(P->n_split, (unsigned)std::min<int64_t>((cols + 255) / 256, 64));
    agg_fixup_kernel<<<g2, 256, 0, st>>>(a, n_out);
    CTB_LAUNCH_CHECK();
  }
  return CTB_OK;
}

}  // namespace

int ctb_pack_transform(int transform, const double* params, int n_params, int n_out, CtbTr* out) {
  std::memset(out, 0, sizeof *out);
  if (n_out < 1 || n_out > CTB_MAX_OUT) {
    ctb_set_error("n_out=%d out of range 1..%d", n_out, CTB_MAX_OUT);
    return CTB_ERR_INVALID;
  }
  auto need = [&](int n) {
    if (n_params == n && (n == 0 || params)) return true;
    ctb_set_error("transform %d with n_out=%d needs %d params, got %d", transform, n_out, n, n_params);
    return false;
  };
  switch (transform) {
    case CTB_TR_IDENTITY:
      if (n_out != 1) { ctb_set_error("IDENTITY has n_out=1"); return CTB_ERR_INVALID; }
      return CTB_OK;
    case CTB_TR_POLY:
      if (!need(1 + n_out)) return CTB_ERR_INVALID;
      out->a[0] = params[0];
      for (int j = 0; j < n_out; ++j) {
        const double p = params[1 + j];
        if (p != (double)(int)p || p < -64 || p > 64) {
          ctb_set_error("POLY power %g is not a small integer", p);
          return CTB_ERR_UNSUPPORTED;
        }
        out->ip[j] = (int)p;
      }
      return CTB_OK;
    case CTB_TR_EDD:
      if (!need(n_out)) return CTB_ERR_INVALID;
      for (int j = 0; j < n_out; ++j) out->a[j] = params[j];
      return CTB_OK;
    case CTB_TR_GDD:
      if (!need(2 * n_out)) return CTB_ERR_INVALID;
      for (int j = 0; j < 2 * n_out; ++j) out->a[j] = params[j];
      return CTB_OK;
  }
  ctb_set_error("transform=%d unsupported", transform);
  return CTB_ERR_INVALID;
}

extern "C" size_t ctb_aggregate_workspace_bytes(const ctb_plan* plan, int64_t T, int n_out) {
  if (!plan || T <= 0 || n_out <= 0) return 0;
  return scratch_bytes(plan, T, n_out);
}

extern "C" int ctb_aggregate(const ctb_plan* P, const void* x0, const void* x1, int dtype,
                             int layout, int64_t stride, const int32_t* time_index, int64_t T,
                             int transform, const double* params, int n_params, int n_out,
                             double* out, int64_t out_ld, void* workspace,
                             size_t workspace_bytes, int variant, void* stream) {
  return aggregate_impl(P, x0, x1, dtype, layout, stride, time_index, T, transform, params, n_params, n_out,
                        nullptr, 0, 1, out, out_ld, workspace, workspace_bytes, variant, stream);
}

// ------------------------------------------------------------ time groups ---
extern "C" int ctb_time_groups_create(const int32_t* group_of_day, int64_t T, int device,
                                      ctb_time_groups** out) {
  if (!out || T < 0 || T >= (1ll << 31) || (T > 0 && !group_of_day)) { ctb_set_error("ctb_time_groups_create: bad argument"); return CTB_ERR_INVALID; }
  *out = nullptr;
  for (int64_t t = 0; t < T; ++t) {
    const int32_t g = group_of_day[t], prev = t ? group_of_day[t - 1] : 0;
    if (g < prev || g > prev + 1 || (t == 0 && g != 0)) {
      ctb_set_error("ctb_time_groups_create: group_of_day must start at 0 and grow by 0 or 1 per day (day %lld: %d after %d)",
                    (long long)t, g, prev);
      return CTB_ERR_INVALID;
    }
  }
  ctb_time_groups* G = new ctb_time_groups();
  G->device = device; G->T = T;
  G->n_groups = T ? group_of_day[T - 1] + 1 : 0;
  std::vector<int32_t> lo(std::max(G->n_groups, 1), 0), hi(std::max(G->n_groups, 1), 0);
  for (int64_t t = T - 1; t >= 0; --t) lo[group_of_day[t]] = (int32_t)t;
  for (int64_t t = 0; t < T; ++t) hi[group_of_day[t]] = (int32_t)t + 1;
  G->gk = 1;
  for (int64_t t0 = 0; t0 < T; t0 += CTB_TB)
    G->gk = std::max(G->gk, group_of_day[std::min<int64_t>(t0 + CTB_TB, T) - 1] - group_of_day[t0] + 1);
  CtbDeviceGuard guard(device);
  auto fail = [&](cudaError_t e) {
    ctb_set_error("ctb_time_groups_create: %s", cudaGetErrorString(e));
    cudaFree(G->d_group); cudaFree(G->d_t_lo); cudaFree(G->d_t_hi);
    delete G;
    return CTB_ERR_CUDA;
  };
  cudaError_t e;
  if (guard.err != cudaSuccess) return fail(guard.err);
  const size_t nT = (size_t)std::max<int64_t>(T, 1), nG = lo.size();
  if ((e = cudaMalloc((void**)&G->d_group, nT * 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&G->d_t_lo, nG * 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&G->d_t_hi, nG * 4)) != cudaSuccess) return fail(e);
  if (T && (e = cudaMemcpy(G->d_group, group_of_day, (size_t)T * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpy(G->d_t_lo, lo.data(), nG * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  if ((e = cudaMemcpy(G->d_t_hi, hi.data(), nG * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
  *out = G;
  return CTB_OK;
}

extern "C" void ctb_time_groups_free(ctb_time_groups* G) {
  if (!G) return;
  CtbDeviceGuard guard(G->device);
  cudaFree(G->d_group); cudaFree(G->d_t_lo); cudaFree(G->d_t_hi);
  delete G;
}

extern "C" int32_t ctb_time_groups_count(const ctb_time_groups* G) { return G ? G->n_groups : 0; }

extern "C" size_t ctb_aggregate_grouped_workspace_bytes(const ctb_plan* plan, const ctb_time_groups* G,
                                                        int n_out) {
  if (!plan || !G || G->T <= 0 || n_out <= 0) return 0;
  return scratch_bytes(plan, G->T, n_out) + gpart_bytes(plan, G, n_out);
}

extern "C" int ctb_aggregate_grouped(const ctb_plan* P, const void* x0, const void* x1, int dtype,
                                     int layout, int64_t stride, const int32_t* time_index, int64_t T,
                                     int transform, const double* params, int n_params, int n_out,
                                     const ctb_time_groups* groups, int64_t t_begin, int flush,
                                     double* out, int64_t out_ld,
                                     void* workspace, size_t workspace_bytes, int variant, void* stream) {
  if (!groups) { ctb_set_error("ctb_aggregate_grouped: null time groups"); return CTB_ERR_INVALID; }
  return aggregate_impl(P, x0, x1, dtype, layout, stride, time_index, T, transform, params, n_params, n_out,
                        groups, t_begin, flush, out, out_ld, workspace, workspace_bytes, variant, stream);
}
