// Internal declarations shared by the translation units of libctb.so.
// Public surface: include/ctb.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "ctb.h"

// ---------------------------------------------------------------- errors ---
void ctb_set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_ctb_launches;

#define CTB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ctb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,    \
                    __LINE__);                                                           \
      return CTB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define CTB_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    g_ctb_launches.fetch_add(1, std::memory_order_relaxed);                              \
    CTB_CUDA(cudaGetLastError());                                                        \
  } while (0)

// ------------------------------------------------------------- constants ---
constexpr int CTB_TB = 32;             // days per staging tile (one warp of lanes)
constexpr int CTB_S = CTB_TB + 1;      // smem row stride in elements: odd => conflict-free
constexpr int CTB_PIECE = 4;           // gridcells per staged piece (16 B of f32)
constexpr int CTB_STAGE_THREADS = 256; // 8 warps: each stages 4 days x 8 pieces per step

// Per-bundle metadata blob, copied to shared memory with one cp.async.bulk:
//   part A: [CtbBlobHeader][n_pieces x int32 piece]
//   part B: [n_seg x CtbSeg][n_ent_pad x double w][n_ent_pad x uint32 off]
//           off = byte offset of the entry's staged cell row in the tile: cell * 33 * elem_bytes
// (every section 16-byte aligned; every segment's entries start at a multiple of 4;
//  off_* are byte offsets from the start of part B)
struct CtbBlobHeader {
  int32_t n_pieces, n_seg;
  int32_t off_seg, off_w, off_loc;  // byte offsets from the start of part B
  int32_t n_ent_pad, bytes_a, bytes_b;   // off_loc = offset of the uint32 `off` array
};
struct CtbSeg {
  int32_t target;  // >= 0: region row of `out`; < 0: ~scratch_slot (region split over bundles)
  uint16_t e0_4;   // first entry / 4
  uint16_t n;      // entries
  double rden;     // 1 / (sum of the region's weights); 1.0 for partial rows.  den == 0 gives
                   // inf: 0 * inf = NaN and x * inf = +-inf, like 0/0 and x/0
};
static_assert(sizeof(CtbBlobHeader) == 32 && sizeof(CtbSeg) == 16, "blob layout");

// fused kernel geometry: CTAs of up to 16 warps, two per SM; every thread stages 4 (16-warp CTAs)
// or 8 (8-warp CTAs) 16-byte loads per batch, so a CTA fills a 128-unit tile in two batches.
// Shared memory is deliberately limited to 164 KB per SM (82 KB per CTA): it is carved out
// of the L1, and the L1 that is left bounds the loads in flight -- measured
// (bench_micro/stage_bw3.py): 4.2 TB/s of staging traffic with <= 164 KB/SM, 2.9 TB/s with
// 196 KB, 2.1 TB/s with 228 KB.
constexpr int CTB_TILE_UNITS = 128;                               // 16-byte units per day in a tile
constexpr int CTB_TILE_BYTES = CTB_TILE_UNITS * 4 * CTB_S * 4;    // 67,584 B of shared memory
constexpr int CTB_CTAS_PER_SM = 2;
constexpr int CTB_SMEM_PER_CTA = 164 * 1024 / CTB_CTAS_PER_SM - 1024 - 512;  // 82,432 B
// metadata blob: part A (header + piece list) and part B (segment table + weights +
// staged-cell indices) are contiguous in global memory and arrive as ONE bulk copy
constexpr int CTB_N_WORK_COUNTERS = 64;
constexpr int CTB_META_A_CAP = 32 + CTB_TILE_UNITS * 4;
constexpr int CTB_META_B_CAP = (CTB_SMEM_PER_CTA - CTB_TILE_BYTES - CTB_META_A_CAP) & ~15;

// -------------------------------------------------------------- the plan ---
struct ctb_plan {
  int device = 0;
  int32_t R = 0;
  int64_t n_rows = 0, nnz = 0, ncell = 0;
  int32_t nlat_phys = 0, nlon_phys = 0;

  // K0 products (device): region-sorted CSR over kept rows
  int32_t* d_row_cell = nullptr;  // [n_rows]  physical flat cell of every weights row
  int32_t* d_row_ptr = nullptr;   // [R+1]
  int32_t* d_col = nullptr;       // [nnz]
  double* d_w = nullptr;          // [nnz]
  double* d_den = nullptr;        // [R]

  // staging bundles (device)
  int32_t n_bundles = 0, n_segments = 0;
  int64_t* d_b_blob_off = nullptr;   // [n_bundles+1] byte offsets into d_blob (16 B aligned)
  int4* d_b_desc = nullptr;          // [n_bundles] {off_lo, off_hi, bytes_a, bytes_b}
  // dynamic unit scheduler of the fused kernel: CTB_N_WORK_COUNTERS device counters used round-robin,
  // one per launch (zeroed on the launch's stream), so that launches of one plan on different
  // streams do not share a counter
  int* d_work_counter = nullptr;
  mutable std::atomic<uint32_t> work_counter_slot{0};
  uint8_t* d_blob = nullptr;         // per-bundle metadata blobs, see CtbBlobHeader
  // regions split over several bundles: out[r] = sum(scratch[slot0..slot1)) / den[r]
  int32_t n_split = 0, n_scratch = 0;
  int32_t* d_split_region = nullptr;  // [n_split]
  int32_t* d_split_slot_ptr = nullptr;  // [n_split+1]

  // compact plans: physical piece of every packed piece, as runs of consecutive pieces
  struct PackRun { int32_t phys_piece, n_pieces, packed_piece; };
  std::vector<PackRun> h_pack_runs;
  int compact = 0;
  int elem_bytes = 4;   // element size the staged-cell byte offsets were built for

  // host mirrors for queries
  std::vector<int32_t> h_row_cell;
  std::vector<double> h_row_w;
  std::vector<double> h_den;
  ctb_plan_info info{};
};

// ------------------------------------------------- gridcell transforms -----
struct CtbTr {
  double a[8];  // thresholds / offset
  int ip[4];    // integer powers (POLY)
};

__device__ __forceinline__ double ctb_ipow(double d, int p) {
  // integer power by squaring; p is warp-uniform
  double r = 1.0, b = d;
  int n = p < 0 ? -p : p;
  while (n) {
    if (n & 1) r *= b;
    b *= b;
    n >>= 1;
  }
  return p < 0 ? 1.0 / r : r;
}

// Snyder exceedance degree days, transformations.py:69-89.
//   tmin < e < tmax :  ((M-e)(pi/2 - asin s) + W cos(asin s)) / pi   with s = (e-M)/W
//                    = W * g(s),  g(s) = (sqrt(1-s^2) - s*acos(s)) / pi,  g(-a) = g(a) + a
//   tmax <= e       :  0          (also when tmax is NaN)
//   tmin >= e       :  M - e      (NaN when tmin is NaN)
// g(a) = v^(3/2) * H(v), v = 1 - a: after factoring the (1-a)^(3/2) behaviour at the end point
// the rest is analytic on [0, 1] (nearest singularity at v = 2), and ONE degree-16 polynomial
// fitted to 80-bit reference values (ctb_edd_coeffs.h, gen_edd_coeffs.py) covers the whole
// range.  Against asin/cos in numpy the scheme agrees to 4e-16 * max(|EDD|, W); it replaces an
// fp64 asin, a sqrt and two divisions by 17 FMAs and one sqrt.  `rW` = 1/W is shared by all
// thresholds of a gridcell-day.
#include "ctb_edd_coeffs.h"

// The coefficients live in constant memory so that every DFMA takes its coefficient as a
// constant-bank / uniform-register operand; as literals each one costs two UMOVs per evaluation
// (half of the instructions of the Snyder kernels).
static __constant__ double ctb_edd_H[CTB_EDD_H_N] = CTB_EDD_H_COEFFS;

__device__ __forceinline__ double ctb_edd_g(double a) {   // a in [0, 1]
  const double v = 1.0 - a;
  const double t = fma(2.0, v, -1.0);
  // even/odd split H(t) = E(t^2) + t*O(t^2): two independent Horner chains of 8 instead of one of
  // 16 (the Snyder kernels stall on this dependency chain: 4 warps per scheduler); same 3.5e-16.
  static_assert(CTB_EDD_H_N == 17, "even/odd split below assumes degree 16");
  const double t2 = t * t;
  double pe = ctb_edd_H[16], po = ctb_edd_H[15];
#pragma unroll
  for (int k = 14; k >= 0; k -= 2) pe = fma(pe, t2, ctb_edd_H[k]);
#pragma unroll
  for (int k = 13; k >= 1; k -= 2) po = fma(po, t2, ctb_edd_H[k]);
  return v * sqrt(v) * fma(po, t, pe);
}

__device__ __forceinline__ double ctb_edd(double tmin, double tmax, double M, double W, double rW,
                                          double e) {
  double r;
  if (tmin < e) {
    if (tmax > e) {
      const double s = (e - M) * rW, a = fabs(s);
      const double g = ctb_edd_g(fmin(a, 1.0));
      r = W * (s < 0.0 ? g + a : g);
    } else {
      r = 0.0;
    }
  } else {
    r = M - e;
  }
  return r;
}

template <int KIND, int NOUT>
__device__ __forceinline__ void ctb_apply(const CtbTr& P, double x0, double x1, double (&f)[NOUT]) {
  if constexpr (KIND == CTB_TR_IDENTITY) {
    f[0] = x0;
  } else if constexpr (KIND == CTB_TR_POLY) {
    const double d = x0 - P.a[0];
    bool chain = true;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) chain = chain && (P.ip[j] == j + 1);
    if (chain) {  // powers 1..NOUT from one read (config 3)
      double p = d;
#pragma unroll
      for (int j = 0; j < NOUT; ++j) {
        f[j] = p;
        p *= d;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NOUT; ++j) f[j] = ctb_ipow(d, P.ip[j]);
    }
  } else if constexpr (KIND == CTB_TR_EDD) {
    const double M = (x1 + x0) / 2, W = (x1 - x0) / 2, rW = 1.0 / W;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) f[j] = ctb_edd(x0, x1, M, W, rW, P.a[j]);
  } else {  // GDD
    const double M = (x1 + x0) / 2, W = (x1 - x0) / 2, rW = 1.0 / W;
#pragma unroll
    for (int j = 0; j < NOUT; ++j)
      f[j] = ctb_edd(x0, x1, M, W, rW, P.a[2 * j]) - ctb_edd(x0, x1, M, W, rW, P.a[2 * j + 1]);
  }
}

static inline int ctb_tr_nin(int kind) { return (kind == CTB_TR_EDD || kind == CTB_TR_GDD) ? 2 : 1; }

// host: validate + pack transform params. returns CTB_OK or error.
int ctb_pack_transform(int transform, const double* params, int n_params, int n_out, CtbTr* out);
