// K1+K2+K3: fused stage / gather / segmented weighted sum with gridcell transforms.
//
// Replaces climate_toolbox/aggregations/aggregations.py:27 (gather) and :75-82
// (sum(w*x)/sum(w) per region) with the transforms of
// climate_toolbox/transformations/transformations.py:69-89,139-141,189 fused in.
//
// Staged kernel (TIME_MAJOR input [T][lat][lon], the BCSD layout):
//   one CTA = one bundle (spatially adjacent regions whose gridcell footprint fits a
//   shared-memory tile) x one block of 32 days.
//   stage : 16-byte coalesced global loads of the footprint's 4-cell pieces for 32
//           day-planes, written TRANSPOSED into smem as a cell-major tile
//           sx[cell][day] (row stride 33 words => conflict-free both ways);
//   gather: one warp per region, lane = day; per CSR entry one conflict-free LDS,
//           fp64 FMA, NaN products skipped; out[r][t] = acc / den[r], 256-byte
//           coalesced stores along time.
// Direct kernel (CELL_MAJOR input [lat][lon][T], or any layout as a fallback):
//   one warp per (region, 32-day tile), lane = day, coalesced along time.
#include <algorithm>
#include <cstring>

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
  const int32_t *b_piece_ptr, *pieces;
  const int64_t* b_blob_off;
  const unsigned char* blob;
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

// ---- load one 4-cell piece of one day-plane ------------------------------
template <typename TIN, bool VEC>
__device__ __forceinline__ void load_piece(const TIN* __restrict__ plane, int piece, int64_t ncell,
                                           TIN (&v)[4]) {
  const int64_t c = (int64_t)piece * CTB_PIECE;
  if constexpr (VEC && sizeof(TIN) == 4) {
    const float4 q = ld_stream_f4(reinterpret_cast<const float*>(plane + c));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if constexpr (VEC) {
    const double2 q0 = ld_stream_d2(reinterpret_cast<const double*>(plane + c));
    const double2 q1 = ld_stream_d2(reinterpret_cast<const double*>(plane + c + 2));
    v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (c + j < ncell) ? __ldg(plane + c + j) : TIN(0);
  }
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <typename TIN> struct StageUnroll { static constexpr int v = sizeof(TIN) == 4 ? 8 : 4; };

// one CSR entry: acc_j += w * f_j(x), NaN products skipped (skipna sum, aggregations.py:78)
template <typename TIN, int KIND, int NOUT>
__device__ __forceinline__ void accumulate(const CtbTr& tr, double w, TIN r0, TIN r1,
                                           double (&acc)[NOUT]) {
  double f[NOUT];
  ctb_apply<KIND, NOUT>(tr, (double)r0, (double)r1, f);
  if constexpr (KIND == CTB_TR_IDENTITY || KIND == CTB_TR_POLY) {
    // f is NaN iff x is NaN: one compare in the storage type gates all outputs
    if (r0 == r0) {
#pragma unroll
      for (int j = 0; j < NOUT; ++j) acc[j] = fma(w, f[j], acc[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NOUT; ++j)
      if (f[j] == f[j]) acc[j] = fma(w, f[j], acc[j]);
  }
}

template <typename TIN, int KIND, int NOUT, bool VEC>
__global__ void __launch_bounds__(CTB_STAGE_THREADS, 3)
agg_staged_kernel(const AggArgs a) {
  constexpr int NIN = NIn<KIND>::v;
  constexpr int S = CTB_S;
  constexpr int UNR = StageUnroll<TIN>::v;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_next;
  __shared__ __align__(8) uint64_t s_bar;

  const int b = blockIdx.x;
  const int t0 = blockIdx.y * CTB_TB;
  const int p0 = a.b_piece_ptr[b];
  const int nP = a.b_piece_ptr[b + 1] - p0;
  const int nCells = nP * CTB_PIECE;
  const int64_t blob0 = a.b_blob_off[b];
  const uint32_t blob_bytes = (uint32_t)(a.b_blob_off[b + 1] - blob0);
  TIN* sx = reinterpret_cast<TIN*>(smem_raw);                     // [NIN][nCells][S]
  int* s_piece = reinterpret_cast<int*>(sx + (size_t)NIN * nCells * S);
  unsigned char* s_blob = reinterpret_cast<unsigned char*>(s_piece) + ((nP * 4 + 15) & ~15);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_next = 0;
    mbar_init(&s_bar, 1);
    // segment table + weights + staged-cell indices: one bulk copy, lands during staging
    bulk_g2s(s_blob, a.blob + blob0, blob_bytes, &s_bar);
  }
  for (int i = tid; i < nP; i += CTB_STAGE_THREADS) s_piece[i] = a.pieces[p0 + i];
  __syncthreads();

  // ---------------- stage: [day][piece] global  ->  [cell][day] shared ----------------
  {
    const int l8 = lane & 7, l4 = lane >> 3;
    const int dl = warp * 4 + l4;  // 8 warps x 4 days = CTB_TB
    const int t = t0 + dl;
    if (t < a.T) {
      const int64_t tp = a.tix ? a.tix[t] : t;
#pragma unroll
      for (int in = 0; in < NIN; ++in) {
        const TIN* plane = reinterpret_cast<const TIN*>(in ? a.x1 : a.x0) + tp * a.stride;
        TIN* sd = sx + (size_t)in * nCells * S + dl;
        for (int pg = l8; pg < nP; pg += 8 * UNR) {
          TIN v[UNR][4];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int q = pg + 8 * u;
            if (q < nP) load_piece<TIN, VEC>(plane, s_piece[q], a.ncell, v[u]);
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int q = pg + 8 * u;
            if (q < nP) {
#pragma unroll
              for (int j = 0; j < 4; ++j) sd[(q * CTB_PIECE + j) * S] = v[u][j];
            }
          }
        }
      }
    }
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);

  // ---------------- gather + segmented weighted sum: warp = region, lane = day --------
  const CtbBlobHeader H = *reinterpret_cast<const CtbBlobHeader*>(s_blob);
  const int4* segs = reinterpret_cast<const int4*>(s_blob + sizeof(CtbBlobHeader));
  const double* W = reinterpret_cast<const double*>(s_blob + H.off_w);
  const uint16_t* LOC = reinterpret_cast<const uint16_t*>(s_blob + H.off_loc);
  const int t = t0 + lane;
  const TIN* sx0 = sx + lane;
  const TIN* sx1 = sx + (size_t)nCells * S + lane;
  for (;;) {
    int s = 0;
    if (lane == 0) s = atomicAdd(&s_next, 1);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= H.n_seg) break;
    const int4 sg = segs[s];  // {target, e0, n, -}
    const int target = sg.x;
    const double den = target >= 0 ? __ldg(a.den + target) : 1.0;  // latency hidden by the loop
    double acc[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) acc[j] = 0.0;
    const int e_full = sg.y + (sg.z & ~3), e_end = sg.y + sg.z;
#pragma unroll 2
    for (int e = sg.y; e < e_full; e += 4) {
      const uint2 lc = *reinterpret_cast<const uint2*>(LOC + e);
      const double2 w01 = *reinterpret_cast<const double2*>(W + e);
      const double2 w23 = *reinterpret_cast<const double2*>(W + e + 2);
      const int l0 = lc.x & 0xffff, l1 = lc.x >> 16, l2 = lc.y & 0xffff, l3 = lc.y >> 16;
      TIN r0[4], r1[4];
      r0[0] = sx0[l0 * S]; r0[1] = sx0[l1 * S]; r0[2] = sx0[l2 * S]; r0[3] = sx0[l3 * S];
      if constexpr (NIN == 2) {
        r1[0] = sx1[l0 * S]; r1[1] = sx1[l1 * S]; r1[2] = sx1[l2 * S]; r1[3] = sx1[l3 * S];
      } else {
        r1[0] = r1[1] = r1[2] = r1[3] = TIN(0);
      }
      accumulate<TIN, KIND, NOUT>(a.tr, w01.x, r0[0], r1[0], acc);
      accumulate<TIN, KIND, NOUT>(a.tr, w01.y, r0[1], r1[1], acc);
      accumulate<TIN, KIND, NOUT>(a.tr, w23.x, r0[2], r1[2], acc);
      accumulate<TIN, KIND, NOUT>(a.tr, w23.y, r0[3], r1[3], acc);
    }
    for (int e = e_full; e < e_end; ++e) {  // ragged tail (< 4 entries)
      const int l = LOC[e];
      TIN r1 = TIN(0);
      if constexpr (NIN == 2) r1 = sx1[l * S];
      accumulate<TIN, KIND, NOUT>(a.tr, W[e], sx0[l * S], r1, acc);
    }
    if (t < a.T) {
      if (target >= 0) {
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          a.out[((size_t)j * a.R + target) * a.out_ld + t] = acc[j] / den;
      } else {
        const int slot = ~target;
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          a.scratch[((size_t)j * a.n_scratch + slot) * a.T + t] = acc[j];
      }
    }
  }
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
template <typename TIN, int KIND, int NOUT>
int launch_staged(const ctb_plan* P, const AggArgs& a, bool vec, cudaStream_t st) {
  constexpr int NIN = NIn<KIND>::v;
  const size_t smem = (size_t)NIN * P->info.max_bundle_cells * CTB_S * sizeof(TIN) +
                      (size_t)P->info.max_meta_bytes;
  int dev_max = 0;
  CTB_CUDA(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, P->device));
  if (smem > (size_t)dev_max) {
    ctb_set_error("staging tile of %zu bytes exceeds the device limit %d: rebuild the plan with "
                  "stage_bytes_per_cell_day=%d", smem, dev_max, (int)(NIN * sizeof(TIN)));
    return CTB_ERR_UNSUPPORTED;
  }
  const dim3 grid(P->n_bundles, (a.T + CTB_TB - 1) / CTB_TB);
  if (P->n_bundles > 0 && a.T > 0) {
    if (vec) {
      auto k = agg_staged_kernel<TIN, KIND, NOUT, true>;
      CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, CTB_STAGE_THREADS, smem, st>>>(a);
    } else {
      auto k = agg_staged_kernel<TIN, KIND, NOUT, false>;
      CTB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, CTB_STAGE_THREADS, smem, st>>>(a);
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
  if (variant == 1) return launch_staged<TIN, KIND, NOUT>(P, a, vec, st);
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
  if (variant == 0) variant = (layout == CTB_LAYOUT_TIME_MAJOR) ? 1 : 2;
  if (variant == 1 && layout != CTB_LAYOUT_TIME_MAJOR) { ctb_set_error("staged variant needs TIME_MAJOR input"); return CTB_ERR_INVALID; }
  if (variant != 1 && variant != 2) { ctb_set_error("variant=%d unsupported", variant); return CTB_ERR_INVALID; }
  const size_t need = variant == 1 ? ctb_aggregate_workspace_bytes(P, T, n_out) : 0;
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
  a.den = P->d_den; a.b_piece_ptr = P->d_b_piece_ptr; a.pieces = P->d_pieces;
  a.b_blob_off = P->d_b_blob_off; a.blob = P->d_blob; a.row_ptr = P->d_row_ptr; a.col = P->d_col;
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
