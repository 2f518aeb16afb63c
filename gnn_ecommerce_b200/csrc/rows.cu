// Rows kernel: Y[r] = sum_e w_e X[src_e] with the fused epilogue for the rows the sweep does not take
// -- at the Cosmetics-Shop shape the 1.6 M user rows (mean in-degree ~3), whose sources are the
// 54 K item rows (L2 / L1 resident). LGConv's `propagate` for those rows (reference call site
// src/lightgcn.py:96) plus the ATen passes the epilogue replaces (src/lightgcn.py:93,97;
// src/train_lightgcn.py:147).
//
// Round 1 staged whole row tiles in shared memory with bulk-async copies and balanced the tile's
// edges over the lanes; it was bound by issued instructions (833 warp-instructions per 8-row tile,
// 56 % of the HBM roofline) and its 227 KB of shared memory left no L1 for the gathered item rows.
// This kernel does as little as possible per row instead:
//   * one sub-warp of L lanes per row, V float4 per lane; the row streams (epilogue operands in,
//     results out) are plain coalesced 128-bit loads / stores with streaming cache hints, issued
//     before the gathers so that their HBM latency overlaps the gather latency;
//   * a row's edges are ONE 8-byte (source, weight) record per lane (rows up to L edges: one load),
//     broadcast inside the sub-warp with shuffles; the gathers of a row are issued back to back;
//   * rows are taken in blocks of 256 consecutive rows (all row-stream traffic of a CTA stays inside
//     a 64 KB window per table) but INSIDE a block in descending-degree order (graph build), so the
//     4 sub-warps of a warp work on rows of equal length: no power-law divergence, no edge balancing
//     (a warp-autonomous variant with 32-row blocks held in registers measured 5 % slower);
//   * no shared-memory tiles: the unified L1 keeps the hot item rows (a few hundred items carry half
//     of the edges), which takes that traffic off the L2 -> SM path.
// Edges of a row are added in CSR (= edge-list) order, like the CPU scatter_add_ of the reference.
#include "epilogue.cuh"

namespace lgc {
namespace {

constexpr int kRowsThreads = 256;
#ifndef LGC_ROWS_L64
#define LGC_ROWS_L64 8
#endif
#ifndef LGC_ROWS_G
#define LGC_ROWS_G 2
#endif

// N edges of the current record set (sub-warp lanes t0 .. t0+N-1): all gathers first, then the FMAs in
// edge order. `n` = edges this sub-warp's row still has (rows of a warp differ by at most a few edges
// after the degree sort: the shorter ones are predicated off).
template <int L, int V, int N>
__device__ __forceinline__ void rows_group(const int2 my, int t0, int n, const float* __restrict__ xl, int ld,
                                           float4 (&acc)[V]) {
  float4 xv[N][V];
  float wv[N];
  bool ok[N];
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const int s = __shfl_sync(0xffffffffu, my.x, t0 + t, L);
    wv[t] = __int_as_float(__shfl_sync(0xffffffffu, my.y, t0 + t, L));
    ok[t] = t0 + t < n && s >= 0;                      // s < 0: source row known to be zero (x_mask)
    if (ok[t]) {
      const float* xr = xl + (size_t)(unsigned)s * ld;
#pragma unroll
      for (int v = 0; v < V; ++v) xv[t][v] = ldg_f4_ptx(xr + 4 * L * v);
    }
  }
#pragma unroll
  for (int t = 0; t < N; ++t)
    if (ok[t]) {
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = fma4_packed(wv[t], xv[t][v], acc[v]);
    }
}

// XM: args.x_mask marks the source rows that are not zero; SPLIT: a block's passes are dealt to `split`
// work items (false: split == 1, the compile-time constant keeps the large-graph code as it was)
#ifndef LGC_ROWS_OCC
#define LGC_ROWS_OCC 4          // resident CTAs per SM the compiler must allow: PLAIN epilogues at ld <= 64
#endif
#ifndef LGC_ROWS_OCC_WIDE
#define LGC_ROWS_OCC_WIDE 3     // ADAM / FWD_FINAL epilogues (4-6 operand rows in registers) and wide rows
#endif
template <int L, int V, int MODE, bool XM, bool SPLIT>
__global__ void __launch_bounds__(kRowsThreads, (MODE == EPI_ADAM || MODE == EPI_FWD_FINAL || V > 2) ? LGC_ROWS_OCC_WIDE : LGC_ROWS_OCC)
k_spmm_rows(const int32_t* __restrict__ rowptr, const int2* __restrict__ rec, const uint8_t* __restrict__ perm,
            const int32_t* __restrict__ blk_cnt, int n_blocks, int split_arg, int num_rows,
            const float* __restrict__ x, EpiArgs args) {
  const int split = SPLIT ? split_arg : 1;
  constexpr int LD = 4 * L * V, NSUB = 32 / L, SUBS = (kRowsThreads / 32) * NSUB, G = LGC_ROWS_G;
  static_assert(kRowsThreads == kRowsBlock, "one thread loads one row's metadata");
  __shared__ int s_rp[2][kRowsBlock + 1];
  __shared__ uint8_t s_perm[2][kRowsBlock];
  __shared__ uint32_t s_mask[2][kRowsBlock / 32];      // addend rows that are not zero (all ones: dense addend)
  const int tid = threadIdx.x, lane = tid & 31, wic = tid >> 5, sw = lane / L, sl = lane % L;
  const int sub = wic * NSUB + sw;
  const float* const xl = x + 4 * sl;
  int buf = 0;
  // work item = (block, part): a block's passes are dealt round-robin to `split` items (1 on large
  // graphs; 2 or 4 when there are too few blocks to fill every CTA slot evenly, e.g. one rank's users)
#pragma unroll 1
  for (int item = blockIdx.x; item < n_blocks * split; item += gridDim.x, buf ^= 1) {
    const int b = item / split, part = item - b * split;
    const int row0 = b * kRowsBlock;
    const int cnt = blk_cnt[b];
    s_rp[buf][tid] = rowptr[min(row0 + tid, num_rows)];
    if (tid == 0) s_rp[buf][kRowsBlock] = rowptr[min(row0 + kRowsBlock, num_rows)];
    s_perm[buf][tid] = perm[row0 + tid];
    if ((MODE == EPI_PLAIN || MODE == EPI_ADAM) && tid < kRowsBlock / 32)
      s_mask[buf][tid] = args.addend_mask ? args.addend_mask[(row0 >> 5) + tid] : 0xffffffffu;
    __syncthreads();                                   // two buffers: one barrier per block is enough
    // the sub-warp's row of pass j0: (row, first CSR entry, degree) and its first L records, one per
    // lane -- fetched one pass ahead so that the record latency is off the critical path. Passes
    // alternate direction (the block's rows are in descending-degree order): every warp gets the
    // same share of long and short rows, which is what the barrier above waits for.
    int r_n = row0, e0_n = 0, deg_n = 0;
    int2 my_n = make_int2(0, 0);
    auto fetch = [&](int j0, int pass) {
      const int j = j0 + ((pass & 1) ? SUBS - 1 - sub : sub);
      r_n = row0; e0_n = 0; deg_n = -1; my_n = make_int2(0, 0);
      if (j < cnt) {
        const int lr = s_perm[buf][j];
        r_n = row0 + lr;
        e0_n = s_rp[buf][lr];
        deg_n = s_rp[buf][lr + 1] - e0_n;
        if (sl < deg_n) {
          my_n = __ldg(rec + e0_n + sl);
          if (XM && !((args.x_mask[my_n.x >> 5] >> (my_n.x & 31)) & 1u)) my_n.x = -1;
        }
      }
    };
    fetch(part * SUBS, 0);
    int pass = 0;
#pragma unroll 1
    for (int j0 = part * SUBS; j0 < cnt; j0 += split * SUBS, ++pass) {
      const int r = r_n, e0 = e0_n, deg = deg_n;      // deg < 0: no row for this sub-warp in this pass
      const bool valid = deg >= 0;
      int2 my = my_n;
      fetch(j0 + split * SUBS, pass + 1);
      const int dmax = __reduce_max_sync(0xffffffffu, deg);
      const size_t off = (size_t)r * LD + 4 * sl;
      EpiPre<MODE, 4> pre[V];
      if (MODE != EPI_FWD_FINAL && valid) {
        bool addend_on = true;
        if (MODE == EPI_PLAIN || MODE == EPI_ADAM) addend_on = (s_mask[buf][(r - row0) >> 5] >> (r & 31)) & 1u;
#pragma unroll
        for (int v = 0; v < V; ++v) epi_preload_w<MODE, 4>(args, off + 4 * L * v, pre[v], addend_on);
      }
      float4 acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int base = 0; base < dmax; base += L) {
        if (base > 0) {
          my = make_int2(0, 0);
          if (base + sl < deg) {
            my = __ldg(rec + e0 + base + sl);
            if (XM && !((args.x_mask[my.x >> 5] >> (my.x & 31)) & 1u)) my.x = -1;
          }
        }
        const int n = deg - base;                      // edges left in this row (<= 0: none)
        const int m = min(dmax - base, L);             // warp-uniform: record lanes in use
#pragma unroll 1
        for (int t0 = 0; t0 < m; t0 += G) {
          if (m - t0 >= G) rows_group<L, V, G>(my, t0, n, xl, LD, acc);
          else rows_group<L, V, 1>(my, t0, n, xl, LD, acc);
        }
      }
      if (valid) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float s4[4] = {acc[v].x, acc[v].y, acc[v].z, acc[v].w};
          if (MODE == EPI_FWD_FINAL) epilogue_w<MODE, 4>(args, off + 4 * L * v, s4);
          else epi_finish_w<MODE, 4>(args, off + 4 * L * v, s4, pre[v]);
        }
      }
    }
  }
}

template <int L, int V, int MODE, bool XM>
int launch_rows_lvmx(const RowPlan* p, const float* x, const EpiArgs& a, const int32_t* rowptr, cudaStream_t st) {
  static int occ_dev[64] = {};
  int dev = 0;
  LGC_CUDA(cudaGetDevice(&dev));
  int occ = (dev >= 0 && dev < 64) ? occ_dev[dev] : 0;
  static int sms_dev[64] = {};
  if (!occ) {
    LGC_CUDA(cudaFuncSetAttribute(k_spmm_rows<L, V, MODE, XM, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 16));
    LGC_CUDA(cudaFuncSetAttribute(k_spmm_rows<L, V, MODE, XM, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 16));   // ~36 KB of shared memory, the rest L1
    LGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmm_rows<L, V, MODE, XM, false>, kRowsThreads, 0));
    if (occ < 1) occ = 1;
    int sms = 0;
    LGC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) { occ_dev[dev] = occ; sms_dev[dev] = sms; }
    else sms_dev[0] = sms;
  }
  const int sms = (dev >= 0 && dev < 64) ? sms_dev[dev] : sms_dev[0];
  const int64_t slots = (int64_t)sms * occ;
  // few blocks (one rank's users): finer work items, when the waves then come out shorter. Rounds of work
  // per CTA slot in units of whole blocks, plus ~8 % per extra part for re-reading the block's metadata
  int split = 1;
  double best = 1e30;
  for (int c = 1; c <= 4; c *= 2) {
    const double cost = (double)ceil_div(p->n_blocks * c, slots) / c * (1.0 + 0.08 * (c - 1));
    if (cost < best - 1e-9) { best = cost; split = c; }
  }
  const int grid = (int)std::min<int64_t>(slots, p->n_blocks * split);
  if (grid <= 0) return LGC_OK;
  ProfScope ps(PROF_LIGHT + (MODE & 3), st);
  if (split == 1)
    k_spmm_rows<L, V, MODE, XM, false><<<grid, kRowsThreads, 0, st>>>(rowptr, p->rec, p->perm, p->blk_cnt,
                                                                      (int)p->n_blocks, 1, (int)p->num_rows, x, a);
  else
    k_spmm_rows<L, V, MODE, XM, true><<<grid, kRowsThreads, 0, st>>>(rowptr, p->rec, p->perm, p->blk_cnt,
                                                                     (int)p->n_blocks, split, (int)p->num_rows, x, a);
  return LGC_OK;
}

template <int L, int V, int MODE>
int launch_rows_lvm(const RowPlan* p, const float* x, const EpiArgs& a, const int32_t* rowptr, cudaStream_t st) {
  if ((MODE == EPI_PLAIN || MODE == EPI_ADAM) && a.x_mask)
    return launch_rows_lvmx<L, V, MODE, (MODE == EPI_PLAIN || MODE == EPI_ADAM)>(p, x, a, rowptr, st);
  return launch_rows_lvmx<L, V, MODE, false>(p, x, a, rowptr, st);
}

template <int L, int V>
int launch_rows_lv(const RowPlan* p, const float* x, EpiMode mode, const EpiArgs& a, const int32_t* rowptr,
                   cudaStream_t st) {
  switch (mode) {
    case EPI_PLAIN: return launch_rows_lvm<L, V, EPI_PLAIN>(p, x, a, rowptr, st);
    case EPI_FWD_INIT: return launch_rows_lvm<L, V, EPI_FWD_INIT>(p, x, a, rowptr, st);
    case EPI_FWD_RMW: return launch_rows_lvm<L, V, EPI_FWD_RMW>(p, x, a, rowptr, st);
    case EPI_ADAM: return launch_rows_lvm<L, V, EPI_ADAM>(p, x, a, rowptr, st);
    case EPI_FWD_FINAL: return launch_rows_lvm<L, V, EPI_FWD_FINAL>(p, x, a, rowptr, st);
  }
  return LGC_ERR_INVALID;
}

}  // namespace

int launch_rows(const lgc_graph* g, const RowPlan* plan, int ld, const float* x, EpiMode mode, const EpiArgs& a,
                cudaStream_t st) {
  if (!g || !plan) return LGC_ERR_INVALID;
  if (plan->n_active == 0) return LGC_OK;              // every row went to the sweep
  int rc = LGC_ERR_UNSUPPORTED;
  // sub-warp geometry per row width: L lanes x V float4 (few lanes per row: more rows per warp instruction)
#define LGC_ROWS_CASE(LDV, LL, LV) case LDV: rc = launch_rows_lv<LL, LV>(plan, x, mode, a, g->rowptr, st); break;
  switch (ld) {
    LGC_ROWS_CASE(64, LGC_ROWS_L64, 16 / LGC_ROWS_L64) LGC_ROWS_CASE(128, 16, 2) LGC_ROWS_CASE(192, 16, 3) LGC_ROWS_CASE(256, 16, 4)
    LGC_ROWS_CASE(32, 8, 1) LGC_ROWS_CASE(96, 8, 3) LGC_ROWS_CASE(160, 8, 5)
    LGC_ROWS_CASE(16, 4, 1) LGC_ROWS_CASE(48, 4, 3) LGC_ROWS_CASE(80, 4, 5)
    default: break;
  }
#undef LGC_ROWS_CASE
  if (rc) return rc;
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

}  // namespace lgc
