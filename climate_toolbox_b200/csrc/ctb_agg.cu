// K1+K2+K3: fused stage / gather / segmented weighted sum with gridcell transforms.
//
// Replaces climate_toolbox/aggregations/aggregations.py:27 (gather) and :75-82
// (sum(w*x)/sum(w) per region) with the transforms of
// climate_toolbox/transformations/transformations.py:69-89,139-141,189 fused in.
//
// Fused kernel (TIME_MAJOR input [T][lat][lon], the BCSD layout): CTAs of 16 warps, two per
// SM, each looping over work units handed out by an atomic counter.  A work unit = one
// bundle (spatially adjacent regions whose gridcell footprint fits a shared-memory tile) x a
// chunk of 4 consecutive 32-day blocks, in chunk-major order so that CTAs running together
// read neighbouring bundles of the same days (shared lines meet in L2).
//   metadata: the bundle's piece list, segment table, weights and staged-cell indices arrive
//           as ONE bulk async copy (cp.async.bulk, TMA 1-D) signalled on an mbarrier, once
//           per unit;
//   stage : 16-byte coalesced global loads of the footprint's 4-cell pieces for 32
//           day-planes, 4 in flight per thread, written TRANSPOSED into smem as a cell-major
//           tile sx[cell][day] (row stride 33 words => conflict-free both ways); a 3-input
//           NaN-propagating max of the magnitudes flags tiles that hold a NaN or infinity;
//   gather: one warp per region, lane = day; per 4 CSR entries three vector LDS of
//           metadata + four conflict-free LDS of data, fp64 FMA; finite tiles run the padded
//           entry range without checks, flagged tiles skip NaN products one by one;
//           out[r][t] = acc / den[r], 256-byte coalesced streaming stores along time.
// Only hardware barriers separate the phases.  The phase times of the two resident CTAs add
// up: a CTA with loads outstanding and a CTA reducing out of shared memory do not overlap on
// one SM (role-split experiment), so each phase is kept short instead.
// Shared memory is capped at 164 KB per SM: it is carved out of the L1, and the L1 that is
// left holds the loads in flight (bench_micro/stage_bw3.py: 4.2 -> 2.1 TB/s at 228 KB).
// Alternatives that were built and measured slower are described in DESIGN.md.
// Direct kernel (CELL_MAJOR input [lat][lon][T], or any layout as a fallback):
//   one warp per (region, 32-day tile), lane = day, coalesced along time.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ctb_internal.cuh"

namespace {

struct AggArgs {
  const void* x0;
  const void* x1;
  int64_t stride;
  const int32_t* tix;
  int T;
  int64_t out_ld;
  int64_t ncell;
  int R;
  double* out;
  double* scratch;
  int n_scratch;
  const double* den;
  const int64_t* b_blob_off;
  const unsigned char* blob;
  int n_bundles;
  int n_items;
  const int4* b_desc;
  int tile_stride;  // bytes between tile stages
  int meta_b_stride;  // bytes between part-B metadata slots
  int n_tb;         // time blocks: ceil(T / 32)
  int chunk_tb;     // time blocks per work unit (a CTA keeps one bundle for a whole unit)
  int* work_counter; // device counter for dynamic unit scheduling (zeroed per launch)
  int dbg;          // CTB_DEBUG bits (perf experiments): 1 skip loads, 2 skip STS, 4 skip gather, 16 role split
  const int32_t *row_ptr, *col;
  const double* w;
  const int32_t *split_region, *split_slot_ptr;
  int n_split;
  CtbTr tr;
};

template <int KIND>
struct NIn { static constexpr int v = (KIND == CTB_TR_EDD || KIND == CTB_TR_GDD) ? 2 : 1; };

// streaming 16-byte load: read once per CTA, keep it out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream_d2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n"
      "@p bra DONE_%=;\n"
      "nanosleep.u32 128;\n"   // back off: polls occupy the MIO queue the LDS/STS need
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// shared-memory load by 32-bit shared-window address
template <typename T>
__device__ __forceinline__ T lds_val(uint32_t addr) {
  T v;
  if constexpr (sizeof(T) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// one CSR entry: acc_j += w * f_j(x), NaN products skipped (skipna sum, aggregations.py:78).
// CHECK = false: the tile was seen to hold no NaN while it was staged (IDENTITY / POLY only,
// where f is NaN iff x is).
template <typename TIN, int KIND, int NOUT, bool CHECK = true>
__device__ __forceinline__ void accumulate(const CtbTr& tr, double w, TIN r0, TIN r1,
                                           double (&acc)[NOUT]) {
  double f[NOUT];
  if constexpr (KIND == CTB_TR_IDENTITY) {
    // the product is NaN iff x is NaN (w is finite, non-zero): zero it in the storage type
    const TIN xs = (!CHECK || r0 == r0) ? r0 : TIN(0);
    acc[0] = fma(w, (double)xs, acc[0]);
  } else if constexpr (KIND == CTB_TR_POLY) {
    ctb_apply<KIND, NOUT>(tr, (double)r0, (double)r1, f);
    if (!CHECK || r0 == r0) {   // f is NaN iff x is NaN: one compare gates all outputs
#pragma unroll
      for (int j = 0; j < NOUT; ++j) acc[j] = fma(w, f[j], acc[j]);
    }
  } else {
    ctb_apply<KIND, NOUT>(tr, (double)r0, (double)r1, f);
#pragma unroll
    for (int j = 0; j < NOUT; ++j)
      if (f[j] == f[j]) acc[j] = fma(w, f[j], acc[j]);
  }
}

template <typename TIN, int KIND, int NOUT, bool VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS >= 1024 ? 1 : 2))
agg_fused_kernel(const AggArgs a) {
  constexpr int NIN = NIn<KIND>::v;
  constexpr int TILE_LOADS = THREADS >= 512 ? 4 : 8;   // loads per thread in flight; 4 suffice with 32 loading warps per SM
  constexpr int S = CTB_S;
  constexpr int HALVES = sizeof(TIN) / 4;       // 16-byte units per piece-day of one input
  constexpr int CPU = 16 / sizeof(TIN);          // cells per unit
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_unit;
  __shared__ int s_nan[2];   // a NaN was staged into the tile (by tile parity)
  __shared__ int s_seg_next;  // Snyder transforms: next region of the tile nobody has taken yet

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* const s_blob = smem_raw + a.tile_stride;
  // experiment (CTB_DEBUG bit 16): first resident CTA of an SM only loads, the second only reduces
  const int dbg = (a.dbg & 16) ? ((blockIdx.x < gridDim.x / 2) ? 6 : 1) : a.dbg;
  if (tid == 0) mbar_init(&s_bar, 1);

  // Work unit = (bundle, chunk of `chunk_tb` consecutive 32-day blocks), handed out by an
  // atomic counter in chunk-major order: CTAs running together work on neighbouring bundles
  // of the same days (shared lines meet in L2), and the bundle's metadata is fetched once
  // per unit.
  for (int n_done = 0;; ++n_done) {
  __syncthreads();   // previous unit fully reduced: tile, blob and s_unit may be reused
  if (tid == 0) {
    s_nan[0] = s_nan[1] = 0;
    s_unit = atomicAdd(a.work_counter, 1);
    if (s_unit < a.n_items) {
      const int4 d = __ldg(a.b_desc + s_unit % a.n_bundles);
      const int64_t o = ((int64_t)(uint32_t)d.y << 32) | (uint32_t)d.x;
      bulk_g2s(s_blob, a.blob + o, (uint32_t)(d.z + d.w), &s_bar);
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
    constexpr int NSUB = (THREADS / 32) / 8;            // warps sharing one 4-day group
    const int sub = warp >> 3;
    const int t = t0 + dl;
    const int nPH = nP * HALVES;
    const int n_units = nPH * NIN;                 // unit index: [input][piece][half]
    if (t < a.T && !(dbg & 1)) {
      const int64_t tp = a.tix ? a.tix[t] : t;
      const TIN* p0 = reinterpret_cast<const TIN*>(a.x0) + tp * a.stride;
      const TIN* p1 = NIN == 2 ? reinterpret_cast<const TIN*>(a.x1) + tp * a.stride : p0;
      // keep the day's base pointers in registers: under the 64-register cap the compiler
      // otherwise rebuilds the 64-bit product in front of every load
      asm volatile("" : "+l"(p0));
      if constexpr (NIN == 2) asm volatile("" : "+l"(p1));
      TIN* sx = reinterpret_cast<TIN*>(smem_raw) + dl;
      for (int g0 = l8 + 8 * sub; g0 < n_units; g0 += 8 * NSUB * TILE_LOADS) {
        int off[TILE_LOADS];
        uint32_t v[TILE_LOADS][4];
#pragma unroll
        for (int u = 0; u < TILE_LOADS; ++u) {      // all index reads first, then all loads
          const int g = g0 + 8 * NSUB * u;
          if (g < n_units) {
            // no runtime division here: it costs ~35 instructions per load in front of the LDG
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
          } else {
            v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0u;   // keeps the NaN scan branch-free
          }
        }
        if constexpr (KIND == CTB_TR_IDENTITY) {
          // any NaN or infinity among the staged values?  (the fast reduction below multiplies
          // padding entries of weight 0 into real cells: exact only for finite data)
          bool nan;
          if constexpr (sizeof(TIN) == 4) {
            // NaN-propagating 3-input max of the magnitudes: two instructions per load
            float m = 0.f;
#pragma unroll
            for (int u = 0; u < TILE_LOADS; ++u)
#pragma unroll
              for (int q = 0; q < 4; q += 2)
                asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(m)
                    : "f"(fabsf(__uint_as_float(v[u][q]))), "f"(fabsf(__uint_as_float(v[u][q + 1]))));
            nan = !(m <= 3.402823466e38f);
          } else {
            nan = false;
#pragma unroll
            for (int u = 0; u < TILE_LOADS; ++u)
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const double f = __longlong_as_double(((long long)v[u][2 * q + 1] << 32) | v[u][2 * q]);
                nan |= !(fabs(f) <= 1.7976931348623157e308);
              }
          }
          if (nan) s_nan[(tb - tb_begin) & 1] = 1;
        }
#pragma unroll
        for (int u = 0; u < TILE_LOADS; ++u) {
          const int g = g0 + 8 * NSUB * u;
          if (g < n_units && !(dbg & 2)) {
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
  const double* W = reinterpret_cast<const double*>(mb + H.off_w);
  const uint32_t* OFF = reinterpret_cast<const uint32_t*>(mb + H.off_loc);   // byte offsets of staged cells
  // 32-bit shared-window addresses: one add per staged value (generic pointers cost two)
  const uint32_t sb0 = smem_u32(smem_raw) + lane * (uint32_t)sizeof(TIN);
  const uint32_t sb1 = sb0 + (uint32_t)(nP * CTB_PIECE * S) * (uint32_t)sizeof(TIN);
  const int t = t0 + lane;
  auto at0 = [&](uint32_t o) { return lds_val<TIN>(sb0 + o); };
  auto at1 = [&](uint32_t o) { return lds_val<TIN>(sb1 + o); };
  const bool tile_nan = s_nan[(tb - tb_begin) & 1] != 0;
  if (tid == 0) s_nan[(tb - tb_begin + 1) & 1] = 0;   // flag of the next tile (nobody reads it now)
  // segments are sorted longest-first: round-robin over the warps is balanced
  // Transforms: the reduction of a region costs hundreds to thousands of cycles, so warps take
  // the next region from a shared counter (list scheduling of the longest-first order) instead
  // of a fixed round-robin share; for the plain aggregation the atomic is not worth it.
  constexpr bool DYN_SEGS = KIND != CTB_TR_IDENTITY;   // measured: poly 1.31 -> 1.29 ms, plain 0.80 -> 0.82 ms
  const int n_seg_run = (dbg & 4) ? 0 : H.n_seg;
  for (int s = warp; s < n_seg_run;) {
    const CtbSeg sg = segs[s];
    double acc[NOUT], acc2[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) acc[j] = acc2[j] = 0.0;
    if constexpr (KIND == CTB_TR_EDD || KIND == CTB_TR_GDD) {
    // Chunks of 4 entries, software-pipelined: the metadata and the staged values of chunk
      // c+1 are loaded before chunk c is accumulated, so every LDS has a full chunk of
      // arithmetic between issue and use.  Entry ranges are padded to a multiple of 4 in the
      // blob (w = 0, cell 0); the padding of the last chunk is masked out.
      const int e0 = (int)sg.e0_4 * 4;
      const int n_chunks = ((int)sg.n + 3) >> 2;
      const int last_valid = (int)sg.n - 4 * (n_chunks - 1);   // 1..4 entries in the last chunk
      double wA[4];
      TIN xA[4], yA[4];
      auto fetch = [&](int e, double (&w)[4], TIN (&x0)[4], TIN (&x1)[4]) {
        const uint4 o = *reinterpret_cast<const uint4*>(OFF + e);
        const double2 w01 = *reinterpret_cast<const double2*>(W + e);
        const double2 w23 = *reinterpret_cast<const double2*>(W + e + 2);
        w[0] = w01.x; w[1] = w01.y; w[2] = w23.x; w[3] = w23.y;
        x0[0] = at0(o.x); x0[1] = at0(o.y); x0[2] = at0(o.z); x0[3] = at0(o.w);
        if constexpr (NIN == 2) {
          x1[0] = at1(o.x); x1[1] = at1(o.y); x1[2] = at1(o.z); x1[3] = at1(o.w);
        } else {
          x1[0] = x1[1] = x1[2] = x1[3] = TIN(0);
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
    } else {
      // plain aggregation / polynomials: 64-register CTAs have no room for the pipelined
      // form (it spills and measured 20 % slower); two chunks per iteration instead.  The
      // metadata LDS feeds the data LDS directly (byte offsets are baked into the plan), and
      // tiles that were staged without a NaN skip the per-value check.
      const int e0 = (int)sg.e0_4 * 4;
      const int e_full = e0 + ((int)sg.n & ~3), e_end = e0 + (int)sg.n;
      auto body = [&](auto check) {
        constexpr bool CHECK = decltype(check)::value;
#pragma unroll 2
        for (int e = e0; e < e_full; e += 4) {
          const uint4 o = *reinterpret_cast<const uint4*>(OFF + e);
          const double2 w01 = *reinterpret_cast<const double2*>(W + e);
          const double2 w23 = *reinterpret_cast<const double2*>(W + e + 2);
          const TIN x0 = at0(o.x), x1 = at0(o.y), x2 = at0(o.z), x3 = at0(o.w);
          accumulate<TIN, KIND, NOUT, CHECK>(a.tr, w01.x, x0, TIN(0), acc);
          accumulate<TIN, KIND, NOUT, CHECK>(a.tr, w01.y, x1, TIN(0), acc2);
          accumulate<TIN, KIND, NOUT, CHECK>(a.tr, w23.x, x2, TIN(0), acc);
          accumulate<TIN, KIND, NOUT, CHECK>(a.tr, w23.y, x3, TIN(0), acc2);
        }
        for (int e = e_full; e < e_end; ++e)   // ragged tail (< 4 entries)
          accumulate<TIN, KIND, NOUT, CHECK>(a.tr, W[e], at0(OFF[e]), TIN(0), acc);
      };
      if constexpr (KIND == CTB_TR_IDENTITY) {
        if (tile_nan) {
          body(std::true_type{});
        } else {
          // all staged values are finite: run the entry range padded to a multiple of 4 (the
          // padding has weight 0 and points at cell 0), no ragged tail, no per-value check
          const int e_pad = e0 + (((int)sg.n + 3) & ~3);
#pragma unroll 2
          for (int e = e0; e < e_pad; e += 4) {
            const uint4 o = *reinterpret_cast<const uint4*>(OFF + e);
            const double2 w01 = *reinterpret_cast<const double2*>(W + e);
            const double2 w23 = *reinterpret_cast<const double2*>(W + e + 2);
            const TIN x0 = at0(o.x), x1 = at0(o.y), x2 = at0(o.z), x3 = at0(o.w);
            acc[0] = fma(w01.x, (double)x0, acc[0]);
            acc2[0] = fma(w01.y, (double)x1, acc2[0]);
            acc[0] = fma(w23.x, (double)x2, acc[0]);
            acc2[0] = fma(w23.y, (double)x3, acc2[0]);
          }
        }
      } else {
        body(std::true_type{});   // polynomials: one code path (two measured slower)
      }
    }
    if (t < a.T) {
      if (sg.target >= 0) {
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          __stcs(&a.out[((size_t)j * a.R + sg.target) * a.out_ld + t], (acc[j] + acc2[j]) * sg.rden);   // written once: leave L2 to the input
      } else {
        const int slot_o = ~sg.target;
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          a.scratch[((size_t)j * a.n_scratch + slot_o) * a.T + t] = acc[j] + acc2[j];
      }
    }
    if constexpr (DYN_SEGS) {
      int nx = 0;
      if (lane == 0) nx = (THREADS / 32) + atomicAdd(&s_seg_next, 1);
      s = __shfl_sync(0xffffffffu, nx, 0);
    } else {
      s += THREADS / 32;
    }
  }
  }   // tiles of the unit
  }   // units
}

// Regions split over several bundles (and regions with no kept rows):
// out[r][t] = (sum of the region's partial rows, fixed order) / den[r].
template <int NOUT_DUMMY>
__global__ void agg_fixup_kernel(const AggArgs a, int n_out) {
  const int i = blockIdx.x;
  const int r = a.split_region[i];
  const int s0 = a.split_slot_ptr[i], s1 = a.split_slot_ptr[i + 1];
  const double d = a.den[r];
  for (int j = 0; j < n_out; ++j)
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < a.T; t += gridDim.y * blockDim.x) {
      double s = 0.0;
      for (int q = s0; q < s1; ++q) s += a.scratch[((size_t)j * a.n_scratch + q) * a.T + t];
      a.out[((size_t)j * a.R + r) * a.out_ld + t] = s / d;
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
    const int t = (int)(wk % n_tiles) * 32 + lane;
    const bool tv = t < a.T;
    const int64_t tp = tv ? (a.tix ? a.tix[t] : t) : 0;
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
    if (tv) {
      const double d = a.den[r];
#pragma unroll
      for (int j = 0; j < NOUT; ++j) a.out[((size_t)j * a.R + r) * a.out_ld + t] = acc[j] / d;
    }
  }
}

// ------------------------------------------------------------- dispatch -----
template <typename TIN, int KIND, int NOUT, int THREADS>
int launch_staged(const ctb_plan* P, AggArgs a, bool vec, cudaStream_t st) {
  constexpr int NIN = NIn<KIND>::v;
  const size_t tile = ((size_t)NIN * P->info.max_bundle_cells * CTB_S * sizeof(TIN) + 127) & ~(size_t)127;
  const size_t meta = (size_t)CTB_META_A_CAP + (((size_t)P->info.max_meta_bytes + 15) & ~(size_t)15);
  const size_t smem = tile + meta;
  if (P->elem_bytes != (int)sizeof(TIN)) {
    ctb_set_error("plan was built for %d-byte elements, input has %d-byte elements: rebuild it with "
                  "elem_bytes=%d", P->elem_bytes, (int)sizeof(TIN), (int)sizeof(TIN));
    return CTB_ERR_UNSUPPORTED;
  }
  int n_sm = 0;
  CTB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, P->device));
  const int by_smem = (int)((164 * 1024) / (smem + 1024 + 64));
  int ctas_per_sm = std::max(1, std::min(THREADS >= 1024 ? 1 : 2, by_smem));
  if (const char* e = getenv("CTB_CTAS")) ctas_per_sm = std::max(1, std::min(ctas_per_sm, atoi(e)));   // experiments
  const int n_tb = (a.T + CTB_TB - 1) / CTB_TB;
  int chunk_tb = 4;
  if (const char* e = getenv("CTB_CHUNK_TB")) chunk_tb = std::max(1, atoi(e));
  const int n_chunks = (n_tb + chunk_tb - 1) / std::max(chunk_tb, 1);
  chunk_tb = n_chunks ? (n_tb + n_chunks - 1) / n_chunks : 1;
  const int64_t n_units = (int64_t)P->n_bundles * n_chunks;
  if (n_units >= (1ll << 31)) { ctb_set_error("too many work units"); return CTB_ERR_UNSUPPORTED; }
  a.n_bundles = P->n_bundles; a.n_items = (int)n_units; a.n_tb = n_tb; a.chunk_tb = std::max(chunk_tb, 1);
  a.tile_stride = (int)tile; a.meta_b_stride = 0; a.work_counter = P->d_work_counter + P->work_counter_slot.fetch_add(1) % CTB_N_WORK_COUNTERS;
  { const char* e = getenv("CTB_DEBUG"); a.dbg = e ? atoi(e) : 0; }
  if (n_units > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(n_units, (int64_t)n_sm * ctas_per_sm);
    const int carveout = (int)((ctas_per_sm * (smem + 1024 + 64) + 2048) * 100 / (228 * 1024)) + 1;
    CTB_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int), st));
    if (vec) {
      auto k = agg_fused_kernel<TIN, KIND, NOUT, true, THREADS>;
      CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(carveout, 100)));
      k<<<grid, THREADS, smem, st>>>(a);
    } else {
      auto k = agg_fused_kernel<TIN, KIND, NOUT, false, THREADS>;
      CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(carveout, 100)));
      k<<<grid, THREADS, smem, st>>>(a);
    }
    CTB_LAUNCH_CHECK();
  }
  if (P->n_split > 0 && a.T > 0) {
    const dim3 g2(P->n_split, (unsigned)std::min<int64_t>((a.T + 255) / 256, 64));
    agg_fixup_kernel<0><<<g2, 256, 0, st>>>(a, NOUT);
    CTB_LAUNCH_CHECK();
  }
  return CTB_OK;
}

template <typename TIN, int KIND, int NOUT>
int launch_direct(const ctb_plan* P, const AggArgs& a, int layout, cudaStream_t st) {
  const int64_t n_work = (int64_t)a.R * ((a.T + 31) / 32);
  if (n_work == 0) return CTB_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((n_work + 7) / 8, 148 * 64);
  if (layout == CTB_LAYOUT_CELL_MAJOR)
    agg_direct_kernel<TIN, KIND, NOUT, CTB_LAYOUT_CELL_MAJOR><<<grid, 256, 0, st>>>(a);
  else
    agg_direct_kernel<TIN, KIND, NOUT, CTB_LAYOUT_TIME_MAJOR><<<grid, 256, 0, st>>>(a);
  CTB_LAUNCH_CHECK();
  (void)P;
  return CTB_OK;
}

template <typename TIN, int KIND, int NOUT>
int run(const ctb_plan* P, const AggArgs& a, int layout, int variant, bool vec, cudaStream_t st) {
  // 16 warps per CTA for the plain aggregation (64 registers suffice); the transform
  // variants take 8 warps with up to 128 registers (no spills in the fp64 EDD/poly code)
  constexpr int THREADS = (KIND == CTB_TR_IDENTITY) ? 512 : 256;
  if (variant == 1) return launch_staged<TIN, KIND, NOUT, THREADS>(P, a, vec, st);
  return launch_direct<TIN, KIND, NOUT>(P, a, layout, st);
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
  return (size_t)plan->n_scratch * (size_t)T * (size_t)n_out * sizeof(double);
}

extern "C" int ctb_aggregate(const ctb_plan* P, const void* x0, const void* x1, int dtype,
                             int layout, int64_t stride, const int32_t* time_index, int64_t T,
                             int transform, const double* params, int n_params, int n_out,
                             double* out, int64_t out_ld, void* workspace,
                             size_t workspace_bytes, int variant, void* stream) {
  if (!P || !x0 || (!out && T > 0 && P->R > 0)) { ctb_set_error("ctb_aggregate: null argument"); return CTB_ERR_INVALID; }
  if (out_ld == 0) out_ld = T;
  if (T < 0 || T >= (1ll << 31) || stride < 0 || out_ld < T) { ctb_set_error("ctb_aggregate: bad T/stride"); return CTB_ERR_INVALID; }
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
  const size_t need = variant != 2 ? ctb_aggregate_workspace_bytes(P, T, n_out) : 0;
  if (need > 0 && (!workspace || workspace_bytes < need)) {
    ctb_set_error("workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    return CTB_ERR_INVALID;
  }
  if (T == 0 || P->R == 0) return CTB_OK;

  int prev = 0;
  CTB_CUDA(cudaGetDevice(&prev));
  if (prev != P->device) CTB_CUDA(cudaSetDevice(P->device));
  a.x0 = x0; a.x1 = x1; a.stride = stride; a.tix = time_index; a.T = (int)T; a.out_ld = out_ld; a.ncell = P->ncell;
  a.R = P->R; a.out = out; a.scratch = (double*)workspace; a.n_scratch = P->n_scratch;
  a.den = P->d_den;
  a.b_blob_off = P->d_b_blob_off; a.blob = P->d_blob; a.b_desc = P->d_b_desc; a.row_ptr = P->d_row_ptr; a.col = P->d_col;
  a.w = P->d_w; a.split_region = P->d_split_region; a.split_slot_ptr = P->d_split_slot_ptr;
  a.n_split = P->n_split;
  const size_t es = dtype == CTB_F32 ? 4 : 8;
  const bool vec = (P->ncell % CTB_PIECE == 0) && ((stride * es) % 16 == 0) &&
                   ((uintptr_t)x0 % 16 == 0) && (!x1 || (uintptr_t)x1 % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  rc = dtype == CTB_F32 ? run_kind<float>(P, a, layout, variant, vec, transform, n_out, st)
                        : run_kind<double>(P, a, layout, variant, vec, transform, n_out, st);
  if (prev != P->device) cudaSetDevice(prev);
  return rc;
}

// ---------------------------------------------------------------- diagnostics ---
// Loads-only replay of the streaming kernel's staging traffic on the plan's real
// footprint (no shared memory, no reduction): measures what the memory system delivers
// for this access pattern as a function of the lane mapping and loads in flight.
namespace {
template <int UNR>
__global__ void debug_stage_bw_kernel(const float* __restrict__ x, int64_t stride, int T,
                                      const int64_t* __restrict__ b_blob_off,
                                      const unsigned char* __restrict__ blob, int n_bundles,
                                      int n_items, int lanes_p, float* sink, int pf_blocks, int order_chunk, int sync_mode) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int lp = lane % lanes_p, ld = lane / lanes_p, dpw = 32 / lanes_p, ndg = 32 / dpw;
  float acc = 0.f;
  for (int64_t idx0 = blockIdx.x; idx0 < n_items; idx0 += gridDim.x) {
    int b = (int)(idx0 % n_bundles), t0 = (int)(idx0 / n_bundles) * CTB_TB;
    if (order_chunk > 0) {
      // unit-major order: a CTA keeps one bundle for `order_chunk` consecutive time blocks
      const int n_tb = n_items / n_bundles, n_ch = (n_tb + order_chunk - 1) / order_chunk;
      const int64_t k = idx0 / gridDim.x;                    // CTA-local item number
      const int64_t unit = blockIdx.x + (k / order_chunk) * gridDim.x;
      const int tb = (int)(unit / n_bundles) * order_chunk + (int)(k % order_chunk);
      if (unit >= (int64_t)n_bundles * n_ch || tb >= n_tb) continue;
      b = (int)(unit % n_bundles); t0 = tb * CTB_TB;
    }
    const unsigned char* bl = blob + b_blob_off[b];
    const int nP = reinterpret_cast<const CtbBlobHeader*>(bl)->n_pieces;
    const int* pieces = reinterpret_cast<const int*>(bl + sizeof(CtbBlobHeader));
    const int chunks = (nP + lanes_p - 1) / lanes_p, units = chunks * ndg;
    for (int u0 = warp; u0 < units; u0 += nw * UNR) {
      float4 v[UNR];
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int u = u0 + k * nw;
        v[k] = make_float4(0, 0, 0, 0);
        if (u < units) {
          const int dg = u % ndg, q = (u / ndg) * lanes_p + lp, t = t0 + dg * dpw + ld;
          if (q < nP && t < T) {
            const float* src = x + (int64_t)t * stride + (int64_t)__ldg(pieces + q) * 4;
            // optional L2 prefetch of the same footprint `pf` time blocks ahead
            if (pf_blocks > 0 && t + pf_blocks * CTB_TB < T)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (int64_t)pf_blocks * CTB_TB * stride));
            v[k] = ld_stream_f4(src);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < UNR; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (sync_mode) __syncthreads();   // CTA-wide barrier per tile, like the fused kernel
  }
  if (acc == 123.25f) *sink = acc;
}
}  // namespace

extern "C" int ctb_debug_stage_bw(const ctb_plan* P, const void* x, int64_t stride, int64_t T,
                                  int lanes_p, int unroll, int warps, int ctas_per_sm, void* sink,
                                  void* stream) {
  if (!P || !x || !sink || (lanes_p != 8 && lanes_p != 16 && lanes_p != 32 && lanes_p != 4)) {
    ctb_set_error("ctb_debug_stage_bw: bad argument");
    return CTB_ERR_INVALID;
  }
  int n_sm = 0;
  CTB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, P->device));
  if (const char* e = getenv("CTB_L2_FETCH")) {
    CTB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    static size_t last = 0;
    if (got != last) { fprintf(stderr, "[ctb] L2 fetch granularity = %zu\n", got); last = got; }
  }
  const int64_t n_items = (int64_t)P->n_bundles * ((T + CTB_TB - 1) / CTB_TB);
  const unsigned grid = (unsigned)std::min<int64_t>(n_items, (int64_t)n_sm * ctas_per_sm);
  cudaStream_t st = (cudaStream_t)stream;
  size_t dsm = 0;
  int pf_blocks = 0;
  if (const char* e = getenv("CTB_DBG_SMEM")) dsm = (size_t)atoi(e);
  if (const char* e = getenv("CTB_DBG_PF")) pf_blocks = atoi(e);
  int order_chunk = 0, sync_mode = 0;
  if (const char* e = getenv("CTB_DBG_SYNC")) sync_mode = atoi(e);
  if (const char* e = getenv("CTB_DBG_ORDER")) order_chunk = atoi(e);
#define CTB_DBG(U) if (dsm) CTB_CUDA(cudaFuncSetAttribute(debug_stage_bw_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm)); debug_stage_bw_kernel<U><<<grid, warps * 32, dsm, st>>>((const float*)x, stride, (int)T, P->d_b_blob_off, P->d_blob, P->n_bundles, (int)n_items, lanes_p, (float*)sink, pf_blocks, order_chunk, sync_mode)
  switch (unroll) {
    case 2: CTB_DBG(2); break;
    case 4: CTB_DBG(4); break;
    case 8: CTB_DBG(8); break;
    case 16: CTB_DBG(16); break;
    default: ctb_set_error("unroll must be 2, 4, 8 or 16"); return CTB_ERR_INVALID;
  }
#undef CTB_DBG
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

// cp.async (LDGSTS) replay of the staging traffic: loader warps copy the footprint of item
// k (pairs of cells x 32 days) straight into a transposed shared-memory tile
// [cell-group][day][width]; `nbuf` tile buffers in flight; a dummy consumer releases them.
namespace {
template <int WIDTH>   // bytes per copy: 4, 8, 16
__global__ void __launch_bounds__(1024, 1)
debug_cpasync_bw_kernel(const float* __restrict__ x, int64_t stride, int T,
                        const int4* __restrict__ b_desc, const unsigned char* __restrict__ blob,
                        int n_bundles, int n_items, int loader_warps, int nbuf, int buf_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_full[8], s_empty[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_loader = loader_warps * 32;
  if (tid == 0) {
    for (int i = 0; i < nbuf; ++i) { mbar_init(&s_full[i], n_loader); mbar_init(&s_empty[i], 1); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int n_my = n_items > (int)blockIdx.x ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr int CPP = 16 / WIDTH;          // copies per 16-byte piece
  constexpr int S = CTB_TB + 1;
  if (warp < loader_warps) {
    for (int k = 0; k < n_my; ++k) {
      const int64_t idx = blockIdx.x + (int64_t)k * gridDim.x;
      const int b = (int)(idx % n_bundles), t0 = (int)(idx / n_bundles) * CTB_TB;
      const int buf = k % nbuf;
      if (k >= nbuf) mbar_wait(&s_empty[buf], ((k / nbuf) - 1) & 1);
      const int4 d = __ldg(b_desc + b);
      const unsigned char* bl = blob + ((((int64_t)(uint32_t)d.y) << 32) | (uint32_t)d.x);
      const int nP = reinterpret_cast<const CtbBlobHeader*>(bl)->n_pieces;
      const int* pieces = reinterpret_cast<const int*>(bl + sizeof(CtbBlobHeader));
      const int n_units = nP * CPP;         // copies per day
      const uint32_t sbase = smem_u32(smem_raw + (size_t)buf * buf_bytes);
      // work = (unit-chunk of 32 lanes, day); warp w takes chunks round-robin
      const int chunks = (n_units + 31) / 32;
      for (int w = warp; w < chunks * CTB_TB; w += loader_warps) {
        const int dday = w % CTB_TB, u = (w / CTB_TB) * 32 + lane, t = t0 + dday;
        if (u < n_units && t < T) {
          const int piece = __ldg(pieces + u / CPP);
          const float* src = x + (int64_t)t * stride + (int64_t)piece * 4 + (u % CPP) * (WIDTH / 4);
          const uint32_t dst = sbase + (uint32_t)((u * S + dday) * WIDTH);
          if constexpr (WIDTH == 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          else if constexpr (WIDTH == 8)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
          else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_full[buf])) : "memory");
    }
  } else if (warp == loader_warps) {
    for (int k = 0; k < n_my; ++k) {       // dummy consumer
      const int buf = k % nbuf;
      mbar_wait(&s_full[buf], (k / nbuf) & 1);
      if (lane == 0) mbar_arrive(&s_empty[buf]);
    }
  }
}
}  // namespace

extern "C" int ctb_debug_cpasync_bw(const ctb_plan* P, const void* x, int64_t stride, int64_t T,
                                    int width, int loader_warps, int nbuf, void* stream) {
  if (!P || !x || nbuf < 1 || nbuf > 8 || loader_warps < 1 || loader_warps > 31) {
    ctb_set_error("ctb_debug_cpasync_bw: bad argument");
    return CTB_ERR_INVALID;
  }
  int n_sm = 0;
  CTB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, P->device));
  const int64_t n_items = (int64_t)P->n_bundles * ((T + CTB_TB - 1) / CTB_TB);
  const int buf_bytes = ((P->info.max_bundle_cells * (CTB_TB + 1) * 4) + 127) & ~127;
  const size_t smem = (size_t)buf_bytes * nbuf;
  if (smem > 227 * 1024) { ctb_set_error("nbuf too large for the plan's tiles (%zu bytes)", smem); return CTB_ERR_INVALID; }
  const unsigned grid = (unsigned)std::min<int64_t>(n_items, n_sm);
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = (loader_warps + 1) * 32;
#define CTB_CPA(W)                                                                              \
  do {                                                                                          \
    CTB_CUDA(cudaFuncSetAttribute(debug_cpasync_bw_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    debug_cpasync_bw_kernel<W><<<grid, threads, smem, st>>>((const float*)x, stride, (int)T, P->d_b_desc, P->d_blob, \
                                                          P->n_bundles, (int)n_items, loader_warps, nbuf, buf_bytes); \
  } while (0)
  if (width == 16) CTB_CPA(16); else if (width == 8) CTB_CPA(8); else if (width == 4) CTB_CPA(4);
  else { ctb_set_error("width must be 4, 8 or 16"); return CTB_ERR_INVALID; }
#undef CTB_CPA
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}
