// Sweep SpMM of the high-degree rows (the item rows of the bipartite graph and the hub users):
// Y[r] = sum_e w_e X[src_e] for the rows whose neighbour lists are long and whose sources are
// spread over the whole (HBM-sized) table -- LGConv's `propagate` for those rows (reference call
// site src/lightgcn.py:96; PyG index_select + mul + scatter_add).
//
// Pulling a hub row chunk by chunk reads every source row once per neighbour that uses it: at the
// Cosmetics-Shop shape a 256-byte user row is needed by ~3 item rows at unrelated times, and L2 does
// not hold the 410 MB user table (round 1: 1.57x the compulsory DRAM reads, latency-bound at 37 % of
// the HBM roofline). Here the OUTPUT rows are pinned instead of the inputs:
//   * every half-warp ("unit", 2 x 32 warps per SM, one persistent CTA per SM) owns <= S output rows
//     (or pieces of a hub row) whose fp32 accumulators live in shared memory for the whole launch
//     (148 x 64 x S rows: 104 K rows at ld = 64);
//   * a unit's edges -- of all its rows together -- are ordered by SOURCE WINDOW (the source range cut
//     into windows of ~64 edges per unit, ~25 MB of table), and all units walk their lists front to
//     back at the same rate (the schedule balances edges per unit and spreads every hub over many
//     units by interleaving, so every list covers the source range uniformly): at any moment the
//     whole GPU reads one or two windows of the source table, the first unit to touch a row pulls it
//     from HBM, the others hit L2. Measured DRAM traffic = 1.1x the table;
//   * inside a window a unit's edges are grouped by accumulator slot: a run of edges of one output
//     row is summed in registers and added to the shared-memory accumulator once, at the record that
//     carries the end-of-run flag (the first version paid one shared-memory read-modify-write per
//     edge and was bound by the LSU pipe: 78 % L1TEX utilisation, short-scoreboard stalls);
//   * per edge: one 128-bit gather per lane (U rows in flight per lane) and the FMAs. No atomics: a
//     slot has one owner, runs are added in window order, edges of a run in edge-list order, split
//     rows are summed in piece order (deterministic);
//   * at the end every unit applies the fused epilogue to its rows (or writes a partial row).
// The schedule depends on S (hence on the row width) and is built on first use.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <queue>
#include <vector>

#include "epilogue.cuh"

namespace lgc {

#ifndef LGC_SWEEP_WARPS
#define LGC_SWEEP_WARPS 32
#endif
#ifndef LGC_SWEEP_U
#define LGC_SWEEP_U 8
#endif
constexpr int kSweepWarps = LGC_SWEEP_WARPS;          // per CTA; one CTA per SM
constexpr int kSweepUnitsPerCta = 2 * kSweepWarps;    // a unit = 16 lanes
constexpr size_t kSweepSmemBudget = 200 * 1024;       // accumulators per CTA
constexpr int kSweepSlotBits = 6;                     // slot ids 0..S (S = trash slot)
constexpr int kSweepMaxSlots = (1 << kSweepSlotBits) - 2;
constexpr int kSweepSrcBits = 31 - kSweepSlotBits;    // record.x = end-of-run flag (bit 31) | slot | source

struct SweepSched {
  int slots = 0;          // S: output rows (or pieces) per unit
  int src_bits = 0;       // record.x = source | slot << src_bits | end-of-run flag << 31
  int n_ctas = 0, n_warps = 0, n_units = 0;
  int64_t n_rows = 0, n_edges = 0, n_iters = 0;
  int64_t n_split_rows = 0, n_partial_slots = 0;
  int window_span = 1, n_windows = 0;   // source rows per window; windows over the gathered table
  // bookkeeping for the roofline accounting (lgc_graph_plan_info)
  int64_t rows_rows = 0, rows_edges = 0, rows_sources = 0, sweep_sources = 0;
  int32_t* warp_ptr = nullptr;   // [n_warps + 1] first iteration of every warp (32 records each)
  int2* rec = nullptr;           // [n_iters * 32] {source | slot << src_bits, weight bits}
  int2* unit_rows = nullptr;     // [n_units * S] {row, partial slot or -1}; row < 0: unused
  int4* split_rows = nullptr;    // [n_split_rows] {row, first partial slot, pieces, 0}
  // row plan of the remaining ("light") rows for rows.cu -- only when the rows kernel is enabled
  RowPlan plan;
};

void sweep_destroy(SweepSched* s) {
  if (!s) return;
  cudaFree(s->warp_ptr); cudaFree(s->rec); cudaFree(s->unit_rows); cudaFree(s->split_rows);
  cudaFree(s->plan.perm); cudaFree(s->plan.blk_cnt); cudaFree(s->plan.rec);
  delete s;
}
size_t sweep_partial_slots(const SweepSched* s) { return s ? (size_t)s->n_partial_slots : 0; }
const RowPlan* sweep_row_plan(const SweepSched* s) { return (s && s->plan.perm) ? &s->plan : nullptr; }
bool sweep_has_rows(const SweepSched* s) { return s && s->n_iters > 0; }
bool rows_kernel_enabled() {
  static const bool on = [] { const char* e = getenv("LGC_ROWS"); return !(e && atoi(e) == 0); }();
  return on;
}
const int4* sweep_split_rows(const SweepSched* s, int64_t* n) {
  *n = s ? s->n_split_rows : 0;
  return s ? s->split_rows : nullptr;
}

namespace {

int sweep_slots(int ld) {
  if (ld <= 0 || ld % 16) return 0;
  const int fpl = ld / 16;
  const bool ok = fpl <= 6 || fpl == 8 || fpl == 10 || fpl == 12 || fpl == 16;
  if (!ok) return 0;
  long s = (long)(kSweepSmemBudget / ((size_t)kSweepUnitsPerCta * ld * 4)) - 1;
  if (s > kSweepMaxSlots) s = kSweepMaxSlots;
  return s >= 2 ? (int)s : 0;
}

// one warp per sweep row: sort key (unit, source window, slot) and the record of every edge
__global__ void k_sweep_keys(int n_rows, const int32_t* __restrict__ r_row, const int32_t* __restrict__ r_off,
                             const int32_t* __restrict__ r_pbase, const int32_t* __restrict__ r_pieces,
                             const int32_t* __restrict__ piece_code, const int32_t* __restrict__ rowptr,
                             const int32_t* __restrict__ src, const float* __restrict__ w, int src_bits,
                             int window_span, int window_bits, unsigned long long* __restrict__ keys,
                             int2* __restrict__ vals) {
  const int i = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n_rows) return;
  const int row = r_row[i], beg = rowptr[row], deg = rowptr[row + 1] - beg;
  const int off = r_off[i], pbase = r_pbase[i], np = r_pieces[i];
  for (int q = lane; q < deg; q += 32) {
    const int code = piece_code[pbase + q % np];       // hub rows: pieces interleaved over the list
    const unsigned long long unit = (unsigned)code >> kSweepSlotBits, slot = (unsigned)code & ((1u << kSweepSlotBits) - 1u);
    const unsigned s = (unsigned)src[beg + q];
    const unsigned long long win = s / (unsigned)window_span;
    keys[off + q] = (((unit << window_bits) | win) << kSweepSlotBits) | slot;
    vals[off + q] = make_int2((int)(s | ((unsigned)slot << src_bits)), __float_as_int(w[beg + q]));
  }
}

// warp-interleaved record stream: iteration `it` of warp w holds 16 records of unit 2w (lanes 0-15)
// and 16 of unit 2w+1 (lanes 16-31); lists are padded with (source 0, weight 0, trash slot). The last
// record of every run of equal slots carries the end-of-run flag.
__global__ void k_sweep_fill(int64_t n_records, int n_warps, const int32_t* __restrict__ warp_ptr,
                             const int32_t* __restrict__ unit_off, const int2* __restrict__ sorted, int src_bits,
                             int trash_code, int2* __restrict__ rec) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_records) return;
  const int it = (int)(t >> 5), lane = (int)(t & 31);
  int lo = 0, hi = n_warps;                            // last warp with warp_ptr[w] <= it
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (warp_ptr[mid] <= it) lo = mid; else hi = mid;
  }
  const int unit = 2 * lo + (lane >> 4);
  const int i = (it - warp_ptr[lo]) * 16 + (lane & 15);
  const int n = unit_off[unit + 1] - unit_off[unit];
  int2 r = make_int2(trash_code, 0);
  if (i < n) {
    r = sorted[unit_off[unit] + i];
    const bool last = i + 1 == n || (((unsigned)sorted[unit_off[unit] + i + 1].x ^ (unsigned)r.x) >> src_bits) != 0u;
    if (last) r.x |= (int)0x80000000u;
  }
  rec[t] = r;
}

// bitmap of the source rows gathered by the rows of one class (in_sweep == which), one warp per row
__global__ void k_mark_sources(int64_t n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                               const uint8_t* __restrict__ in_sweep, int which, uint32_t* __restrict__ bits) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows || in_sweep[r] != which) return;
  for (int j = rowptr[r] + lane; j < rowptr[r + 1]; j += 32) atomicOr(&bits[src[j] >> 5], 1u << (src[j] & 31));
}
__global__ void k_popcount(int64_t n_words, const uint32_t* __restrict__ bits, unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x)
    c += __popc(bits[i]);
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// (source, weight) of every CSR entry side by side: the rows kernel reads a row's edges as one
// 8-byte record per lane
__global__ void k_interleave(int64_t nnz, const int32_t* __restrict__ src, const float* __restrict__ w, int2* __restrict__ rec) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * blockDim.x)
    rec[j] = make_int2(src[j], __float_as_int(w[j]));
}

template <typename T>
struct Dev {
  T* p = nullptr;
  ~Dev() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  T* release() { T* q = p; p = nullptr; return q; }
};

#define SWEEP_CUDA(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      set_error(std::string("sweep schedule: ") + #expr + ": " + cudaGetErrorString(_e));         \
      return nullptr;                                                                             \
    }                                                                                             \
  } while (0)

SweepSched* sweep_build(const lgc_graph* g, int S) {
  NvtxRange nvtx("lgc_sweep_schedule_build");
  int dev = 0, n_sms = 0;
  SWEEP_CUDA(cudaGetDevice(&dev));
  SWEEP_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_warps = n_sms * kSweepWarps, n_units = 2 * n_warps;
  const int src_bits = kSweepSrcBits;
  if (g->num_cols >= (1LL << src_bits) || S > kSweepMaxSlots) return nullptr;   // all-ones source = "skip" marker

  const int64_t n = g->num_nodes;
  std::vector<int32_t> rp((size_t)n + 1);
  SWEEP_CUDA(cudaDeviceSynchronize());
  SWEEP_CUDA(cudaMemcpy(rp.data(), g->rowptr, ((size_t)n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));

  // ---- sweep rows: by degree, everything above the threshold up to the accumulator capacity. With
  // the rows kernel (rows.cu) the rest may have any degree; with the round-1 light-row kernel
  // (LGC_ROWS=0) every row above its limit must fit.
  const bool with_plan = rows_kernel_enabled();
  static const int rows_max_degree = [] { const char* e = getenv("LGC_ROWS_MAX_DEGREE"); int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();
  // A rectangular operator whose rows all fit (the item-partial operator of the multi-GPU step: every
  // item row, few edges each per rank) goes to the sweep entirely: one launch instead of two, and the
  // rows with few LOCAL edges still share the window of the source table with the hubs.
  const bool all_rows = with_plan && g->num_cols != g->num_nodes && n <= (int64_t)n_units * S * 9 / 10;
  const int threshold = all_rows ? -1 : (with_plan ? rows_max_degree : g->light_max_degree);
  std::vector<int32_t> rows;
  for (int64_t r = 0; r < n; ++r)
    if (rp[r + 1] - rp[r] > threshold) rows.push_back((int32_t)r);
  std::stable_sort(rows.begin(), rows.end(), [&](int32_t a, int32_t b) {
    return rp[a + 1] - rp[a] > rp[b + 1] - rp[b];
  });
  const int64_t capacity = (int64_t)n_units * S * 9 / 10;        // leave slots for the pieces of split rows
  if ((int64_t)rows.size() > capacity) {
    if (!with_plan) return nullptr;
    rows.resize((size_t)capacity);
  }
  if (rows.empty() && !with_plan) return nullptr;
  int64_t n_edges = 0;
  for (int32_t r : rows) n_edges += rp[r + 1] - rp[r];

  // ---- pieces: a row longer than piece_max is cut into equal interleaved pieces
  const int64_t target = std::max<int64_t>(1, (n_edges + n_units - 1) / n_units);      // edges per unit
  int64_t piece_max = std::max<int64_t>(32, target / 3);
  std::vector<int32_t> r_pieces(rows.size()), r_pbase(rows.size()), r_off(rows.size() + 1);
  int64_t n_pieces = 0;
  for (;;) {
    n_pieces = 0;
    for (size_t i = 0; i < rows.size(); ++i) {
      const int d = rp[rows[i] + 1] - rp[rows[i]];
      r_pieces[i] = (int32_t)std::max<int64_t>(1, (d + piece_max - 1) / piece_max);   // a row without edges still owns a slot (its epilogue)
      n_pieces += r_pieces[i];
    }
    if (n_pieces <= (int64_t)n_units * S) break;
    piece_max *= 2;
  }
  struct Piece { int32_t count, ridx, k; };
  std::vector<Piece> pieces;
  pieces.reserve((size_t)n_pieces);
  {
    int64_t off = 0, pb = 0;
    for (size_t i = 0; i < rows.size(); ++i) {
      const int d = rp[rows[i] + 1] - rp[rows[i]], np = r_pieces[i];
      r_off[i] = (int32_t)off; r_pbase[i] = (int32_t)pb;
      for (int k = 0; k < np; ++k) pieces.push_back({d / np + (k < d % np ? 1 : 0), (int32_t)i, k});
      off += d; pb += np;
    }
    r_off[rows.size()] = (int32_t)off;
  }
  std::vector<int32_t> order(pieces.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return pieces[a].count > pieces[b].count; });

  // ---- longest-processing-time assignment of the pieces to the units, <= S pieces per unit
  typedef std::pair<int64_t, int32_t> Load;   // (edges so far, unit)
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int u = 0; u < n_units; ++u) heap.push({0, u});
  std::vector<int32_t> unit_load(n_units, 0), unit_np(n_units, 0), piece_code(pieces.size());
  std::vector<int2> unit_rows((size_t)n_units * S, make_int2(-1, -1));
  std::vector<int4> split_rows;
  std::vector<int32_t> slot_base(rows.size(), -1);
  int64_t n_partial = 0;
  for (size_t i = 0; i < rows.size(); ++i)
    if (r_pieces[i] > 1) {
      slot_base[i] = (int32_t)n_partial;
      split_rows.push_back(make_int4(rows[i], (int)n_partial, r_pieces[i], 0));
      n_partial += r_pieces[i];
    }
  for (int32_t pi : order) {
    if (heap.empty()) return nullptr;
    const Load top = heap.top();
    heap.pop();
    const int u = top.second, slot = unit_np[u]++;
    const Piece& p = pieces[pi];
    piece_code[r_pbase[p.ridx] + p.k] = (u << kSweepSlotBits) + slot;
    unit_rows[(size_t)u * S + slot] = make_int2(rows[p.ridx], slot_base[p.ridx] >= 0 ? slot_base[p.ridx] + p.k : -1);
    unit_load[u] += p.count;
    if (unit_np[u] < S) heap.push({top.first + p.count, u});
  }
  std::vector<int32_t> unit_off(n_units + 1, 0), warp_ptr(n_warps + 1, 0);
  for (int u = 0; u < n_units; ++u) unit_off[u + 1] = unit_off[u] + unit_load[u];
  for (int w = 0; w < n_warps; ++w)
    warp_ptr[w + 1] = warp_ptr[w] + (std::max(unit_load[2 * w], unit_load[2 * w + 1]) + 15) / 16;
  const int64_t n_iters = warp_ptr[n_warps];

  // ---- device side: records sorted by (unit, source), then interleaved per warp
  Dev<int32_t> d_rows, d_off, d_pbase, d_pieces, d_code, d_unit_off, d_warp_ptr;
  Dev<unsigned long long> keys_in, keys_out;
  Dev<int2> vals_in, vals_out, d_rec, d_unit_rows;
  Dev<int4> d_split;
  Dev<char> tmp;
  SWEEP_CUDA(d_rows.upload(rows)); SWEEP_CUDA(d_off.upload(r_off)); SWEEP_CUDA(d_pbase.upload(r_pbase));
  SWEEP_CUDA(d_pieces.upload(r_pieces)); SWEEP_CUDA(d_code.upload(piece_code));
  SWEEP_CUDA(d_unit_off.upload(unit_off)); SWEEP_CUDA(d_warp_ptr.upload(warp_ptr));
  SWEEP_CUDA(d_unit_rows.upload(unit_rows)); SWEEP_CUDA(d_split.upload(split_rows));
  SWEEP_CUDA(keys_in.alloc(n_edges)); SWEEP_CUDA(keys_out.alloc(n_edges));
  SWEEP_CUDA(vals_in.alloc(n_edges)); SWEEP_CUDA(vals_out.alloc(n_edges));
  SWEEP_CUDA(d_rec.alloc((size_t)n_iters * 32));
  const int threads = 256;
  // source windows: ~window_edges edges of every unit per window (LGC_SWEEP_WINDOW, default 64)
  static const int window_edges = [] { const char* e = getenv("LGC_SWEEP_WINDOW"); int v = e ? atoi(e) : 0; return v > 0 ? v : 64; }();
  int64_t n_windows = std::min<int64_t>(4096, std::max<int64_t>(1, (target + window_edges - 1) / window_edges));
  const int window_span = (int)std::max<int64_t>(1, (g->num_cols + n_windows - 1) / n_windows);
  n_windows = (g->num_cols + window_span - 1) / window_span;
  int window_bits = 0;
  while ((1LL << window_bits) < n_windows) ++window_bits;
  if (!rows.empty()) {
    k_sweep_keys<<<(int)ceil_div((int64_t)rows.size() * 32, threads), threads>>>(
        (int)rows.size(), d_rows.p, d_off.p, d_pbase.p, d_pieces.p, d_code.p, g->rowptr, g->src, g->w_hat, src_bits,
        window_span, window_bits, keys_in.p, vals_in.p);
    SWEEP_CUDA(cudaGetLastError());
  }
  int unit_bits = 1;
  while ((1 << unit_bits) < n_units) ++unit_bits;
  const int key_bits = unit_bits + window_bits + kSweepSlotBits;
  size_t tmp_bytes = 0;                                 // stable: equal keys keep the edge-list order
  SWEEP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, vals_out.p,
                                             (int)n_edges, 0, key_bits));
  SWEEP_CUDA(tmp.alloc(tmp_bytes));
  if (n_edges > 0)
    SWEEP_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, vals_out.p,
                                               (int)n_edges, 0, key_bits));
  const int64_t n_records = n_iters * 32;
  if (n_records > 0) {
    k_sweep_fill<<<(int)ceil_div(n_records, threads), threads>>>(n_records, n_warps, d_warp_ptr.p, d_unit_off.p,
                                                                 vals_out.p, src_bits, (int)((unsigned)S << src_bits), d_rec.p);
    SWEEP_CUDA(cudaGetLastError());
  }
  // ---- row plan of the other rows (rows.cu): blocks of kRowsBlock consecutive rows, inside a block
  // the rows ordered by degree (descending) so that the sub-warps of a warp get rows of equal length
  Dev<uint8_t> d_perm;
  Dev<int32_t> d_cnt;
  Dev<int2> d_rec2;
  int64_t n_blocks = 0;
  if (with_plan) {
    n_blocks = (n + kRowsBlock - 1) / kRowsBlock;
    std::vector<uint8_t> in_sweep((size_t)n, 0);
    for (int32_t r : rows) in_sweep[r] = 1;
    std::vector<uint8_t> perm((size_t)n_blocks * kRowsBlock, 0);
    std::vector<int32_t> cnt((size_t)n_blocks, 0);
    std::vector<int32_t> bucket_pos(kRowsBlock);
    for (int64_t b = 0; b < n_blocks; ++b) {
      const int64_t r0 = b * kRowsBlock, r1 = std::min<int64_t>(n, r0 + kRowsBlock);
      int c = 0;
      for (int64_t r = r0; r < r1; ++r)
        if (!in_sweep[r]) bucket_pos[c++] = (int32_t)(r - r0);
      std::stable_sort(bucket_pos.begin(), bucket_pos.begin() + c, [&](int32_t a, int32_t bb) {
        return rp[r0 + a + 1] - rp[r0 + a] > rp[r0 + bb + 1] - rp[r0 + bb];
      });
      for (int i = 0; i < c; ++i) perm[(size_t)r0 + i] = (uint8_t)bucket_pos[i];
      cnt[b] = c;
    }
    SWEEP_CUDA(d_perm.upload(perm)); SWEEP_CUDA(d_cnt.upload(cnt));
    SWEEP_CUDA(d_rec2.alloc((size_t)g->nnz + 8));
    if (g->nnz > 0) {
      k_interleave<<<(int)std::min<int64_t>(ceil_div(g->nnz, threads), 148 * 16), threads>>>(g->nnz, g->src, g->w_hat, d_rec2.p);
      SWEEP_CUDA(cudaGetLastError());
    }
  }
  // ---- distinct gathered source rows per kernel class (algorithmic bytes of a launch)
  unsigned long long h_distinct[2] = {0, 0};
  {
    std::vector<uint8_t> in_sweep((size_t)n, 0);
    for (int32_t r : rows) in_sweep[r] = 1;
    Dev<uint8_t> d_in;
    Dev<uint32_t> d_bits;
    Dev<unsigned long long> d_cnt2;
    const int64_t n_words = (g->num_cols + 31) / 32;
    SWEEP_CUDA(d_in.upload(in_sweep)); SWEEP_CUDA(d_bits.alloc(n_words)); SWEEP_CUDA(d_cnt2.alloc(2));
    SWEEP_CUDA(cudaMemset(d_cnt2.p, 0, 16));
    for (int which = 0; which < 2; ++which) {
      SWEEP_CUDA(cudaMemset(d_bits.p, 0, n_words * 4));
      k_mark_sources<<<(int)ceil_div(n * 32, threads), threads>>>(n, g->rowptr, g->src, d_in.p, which, d_bits.p);
      k_popcount<<<148 * 4, threads>>>(n_words, d_bits.p, d_cnt2.p + which);
      SWEEP_CUDA(cudaGetLastError());
    }
    SWEEP_CUDA(cudaMemcpy(h_distinct, d_cnt2.p, 16, cudaMemcpyDeviceToHost));
  }
  SWEEP_CUDA(cudaDeviceSynchronize());

  SweepSched* s = new SweepSched();
  s->rows_rows = n - (int64_t)rows.size(); s->rows_edges = g->nnz - n_edges;
  s->rows_sources = (int64_t)h_distinct[0]; s->sweep_sources = (int64_t)h_distinct[1];
  s->plan.n_blocks = n_blocks; s->plan.num_rows = n; s->plan.n_active = n - (int64_t)rows.size();
  s->plan.perm = d_perm.release(); s->plan.blk_cnt = d_cnt.release(); s->plan.rec = d_rec2.release();
  s->slots = S; s->src_bits = src_bits;
  s->n_ctas = n_sms; s->n_warps = n_warps; s->n_units = n_units;
  s->n_rows = (int64_t)rows.size(); s->n_edges = n_edges; s->n_iters = n_iters;
  s->n_split_rows = (int64_t)split_rows.size(); s->n_partial_slots = n_partial;
  s->window_span = window_span; s->n_windows = (int)n_windows;
  s->warp_ptr = d_warp_ptr.release(); s->rec = d_rec.release();
  s->unit_rows = d_unit_rows.release(); s->split_rows = d_split.release();
  return s;
}

template <int FPL>
struct SweepCfg {
  static constexpr int LD = 16 * FPL;
  static constexpr int W = FPL % 4 == 0 ? 4 : (FPL % 2 == 0 ? 2 : 1);   // floats per vector access
  static constexpr int NV = FPL / W;                                     // vectors per lane and row
  // gathered rows in flight per lane (one batch): U * FPL data registers
  static constexpr int U0 = FPL <= 4 ? LGC_SWEEP_U : (FPL <= 8 ? LGC_SWEEP_U / 2 : LGC_SWEEP_U / 4);
  static constexpr int U = U0 < 2 ? 2 : U0;
};

// acc += w * x over a lane's FPL floats; packed FFMA2 (same roundings as fmaf, half the issue slots)
template <int FPL>
__device__ __forceinline__ void fma_row(float w, const float (&x)[FPL], float (&acc)[FPL]) {
  if constexpr (FPL % 2 == 0) {
    unsigned long long ww;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
#pragma unroll
    for (int i = 0; i < FPL; i += 2) {
      unsigned long long xx, aa;
      asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x[i]), "f"(x[i + 1]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(acc[i]), "f"(acc[i + 1]));
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(aa) : "l"(ww), "l"(xx));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(aa));
    }
  } else {
#pragma unroll
    for (int i = 0; i < FPL; ++i) acc[i] = fmaf(w, x[i], acc[i]);
  }
}

template <int FPL, int MODE, bool XM>   // XM: args.x_mask marks the source rows that are not zero
__global__ void __launch_bounds__(32 * kSweepWarps, 1)
k_spmm_sweep(const int32_t* __restrict__ warp_ptr, const int2* __restrict__ rec, const int2* __restrict__ unit_rows,
             int S, int src_bits, const float* __restrict__ x, float* __restrict__ partials, EpiArgs args) {
  using C = SweepCfg<FPL>;
  constexpr int LD = C::LD, W = C::W, NV = C::NV, U = C::U;
  extern __shared__ __align__(16) float s_acc[];
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int warp = blockIdx.x * kSweepWarps + wic;
  // vector k of this lane covers columns W * (l16 + 16 k) .. + W - 1 of the unit's rows
  float* const my = s_acc + (size_t)((wic * 2 + half) * (S + 1)) * LD + W * l16;
  {
    float z[W];
#pragma unroll
    for (int i = 0; i < W; ++i) z[i] = 0.f;
    for (int s = 0; s <= S; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) stv<W>(my + s * LD + 16 * W * k, z);
  }
  __syncwarp();

  const int it0 = warp_ptr[warp], it1 = warp_ptr[warp + 1];
  const unsigned src_mask = (1u << src_bits) - 1u;
  const float* const xl = x + W * l16;
  float acc[FPL];                                      // sum of the current run of one output row
#pragma unroll
  for (int i = 0; i < FPL; ++i) acc[i] = 0.f;
  // Tried and dropped (profiles/r2/sweep_notes.md): L2 prefetch of the gathered rows a few iterations
  // ahead (prefetch.global.L2 = CCTL.E.PF2: 20-50 % slower, more with distance) and a sequential
  // bulk prefetch of the next source window by the TMA engine (cp.async.bulk.prefetch.L2: 15-30 %
  // slower). With the table L2-resident the kernel runs only 20 % faster (0.091 vs 0.114 ms for the
  // item rows at c2): it moves 1.1 GB of gathered rows from L2 to the SMs at ~12 TB/s, the measured
  // L2 -> SM ceiling, so hiding the HBM misses has little left to win.
  // records of one iteration, one per lane; a source row that is known to be zero (x_mask, the
  // sparse gradient table of the first backward layer) is marked with the all-ones source id
  const uint32_t* const x_mask = args.x_mask;
  auto load_rec = [&](int i) {
    int2 q = __ldg(rec + (size_t)i * 32 + lane);
    if (XM) {
      const unsigned s = (unsigned)q.x & src_mask;
      if (!((x_mask[s >> 5] >> (s & 31)) & 1u)) q.x |= (int)src_mask;
    }
    return q;
  };
  int2 r = make_int2(0, 0);
  if (it0 < it1) r = load_rec(it0);
  const int base_lane = lane & 16;
#pragma unroll 1
  for (int it = it0; it < it1; ++it) {
    int2 rn = r;
    if (it + 1 < it1) rn = load_rec(it + 1);             // next iteration's records
#pragma unroll
    for (int b = 0; b < 16; b += U) {
      float xv[U][FPL];
      unsigned ev[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ev[u] = (unsigned)__shfl_sync(0xffffffffu, r.x, base_lane + b + u);
        const unsigned sidx = ev[u] & src_mask;
        if (!XM || sidx != src_mask) {
          const float* xr = xl + (size_t)sidx * LD;
#pragma unroll
          for (int k = 0; k < NV; ++k) ldv_nc<W>(xr + 16 * W * k, &xv[u][W * k]);
        } else {
#pragma unroll
          for (int i = 0; i < FPL; ++i) xv[u][i] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float wv = __int_as_float(__shfl_sync(0xffffffffu, r.y, base_lane + b + u));
        fma_row<FPL>(wv, xv[u], acc);
        if ((int)ev[u] < 0) {                          // end of the run: add it to the row's accumulator
          float* a = my + ((ev[u] & 0x7fffffffu) >> src_bits) * LD;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            float t[W];
            ldv<W>(a + 16 * W * k, t);
#pragma unroll
            for (int i = 0; i < W; ++i) { t[i] += acc[W * k + i]; acc[W * k + i] = 0.f; }
            stv<W>(a + 16 * W * k, t);
          }
        }
      }
    }
    r = rn;
  }
  __syncwarp();

  // ---- epilogue: the unit's rows leave shared memory through the fused epilogue (or as partial rows)
  // (the unit's row table is read 16 entries at a time, one per lane, and handed round with shuffles:
  // S dependent global loads in a row made this tail as long as the sweep itself on small operators)
  const int2* ur = unit_rows + (size_t)(warp * 2 + half) * S;
  for (int s0 = 0; s0 < S; s0 += 16) {
    int2 q = make_int2(-1, -1);
    if (s0 + l16 < S) q = __ldg(ur + s0 + l16);
    const int n_here = min(16, S - s0);
    for (int j = 0; j < n_here; ++j) {
      int2 d;
      d.x = __shfl_sync(0xffffffffu, q.x, (lane & 16) + j);
      d.y = __shfl_sync(0xffffffffu, q.y, (lane & 16) + j);
      if (d.x < 0) continue;
      const int s = s0 + j;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float t[W];
        ldv<W>(my + s * LD + 16 * W * k, t);
        const int col = W * l16 + 16 * W * k;
        if (d.y >= 0) stv<W>(partials + (size_t)d.y * LD + col, t);
        else epilogue_w<MODE, W>(args, (size_t)d.x * LD + col, t);
      }
    }
  }
}

template <int FPL, int MODE, bool XM>
int launch_sweep_fmx(const SweepSched* s, const float* x, const EpiArgs& a, float* partials, cudaStream_t st) {
  constexpr int LD = 16 * FPL;
  const size_t smem = (size_t)kSweepUnitsPerCta * (s->slots + 1) * LD * 4;
  static bool attr_set[64] = {};
  int dev = 0;
  LGC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    LGC_CUDA(cudaFuncSetAttribute(k_spmm_sweep<FPL, MODE, XM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSweepSmemBudget));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  ProfScope ps(PROF_HEAVY + (MODE & 3), st);
  k_spmm_sweep<FPL, MODE, XM><<<s->n_ctas, 32 * kSweepWarps, smem, st>>>(s->warp_ptr, s->rec, s->unit_rows, s->slots,
                                                                          s->src_bits, x, partials, a);
  return LGC_OK;
}
template <int FPL, int MODE>
int launch_sweep_fm(const SweepSched* s, const float* x, const EpiArgs& a, float* partials, cudaStream_t st) {
  // the source-mask variant exists for the two epilogues the first backward layer can have
  if ((MODE == EPI_PLAIN || MODE == EPI_ADAM) && a.x_mask)
    return launch_sweep_fmx<FPL, MODE, (MODE == EPI_PLAIN || MODE == EPI_ADAM)>(s, x, a, partials, st);
  return launch_sweep_fmx<FPL, MODE, false>(s, x, a, partials, st);
}

template <int FPL>
int launch_sweep_f(const SweepSched* s, const float* x, EpiMode mode, const EpiArgs& a, float* partials,
                   cudaStream_t st) {
  switch (mode) {
    case EPI_PLAIN: return launch_sweep_fm<FPL, EPI_PLAIN>(s, x, a, partials, st);
    case EPI_FWD_INIT: return launch_sweep_fm<FPL, EPI_FWD_INIT>(s, x, a, partials, st);
    case EPI_FWD_RMW: return launch_sweep_fm<FPL, EPI_FWD_RMW>(s, x, a, partials, st);
    case EPI_ADAM: return launch_sweep_fm<FPL, EPI_ADAM>(s, x, a, partials, st);
    case EPI_FWD_FINAL: return launch_sweep_fm<FPL, EPI_FWD_FINAL>(s, x, a, partials, st);
  }
  return LGC_ERR_INVALID;
}

}  // namespace

const SweepSched* sweep_get(const lgc_graph* g, int ld) {
  static const bool disabled = [] { const char* e = getenv("LGC_SWEEP"); return e && atoi(e) == 0; }();
  if (disabled || !g) return nullptr;
  const int S = sweep_slots(ld);
  if (!S) return nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < kMaxSweepScheds; ++i)
    if (g->sweep[i] && g->sweep[i]->slots == S) return g->sweep[i];
  for (int i = 0; i < g->n_sweep_failed; ++i)
    if (g->sweep_failed[i] == S) return nullptr;
  SweepSched* s = sweep_build(g, S);
  if (s) {
    for (int i = 0; i < kMaxSweepScheds; ++i)
      if (!g->sweep[i]) { g->sweep[i] = s; return s; }
    sweep_destroy(s);            // more distinct row widths than cache entries: not expected
    return nullptr;
  }
  if (g->n_sweep_failed < kMaxSweepScheds) g->sweep_failed[g->n_sweep_failed++] = S;
  return nullptr;
}

void sweep_info(const SweepSched* s, lgc_plan_info* info) {
  info->sweep_rows = s->n_rows; info->sweep_edges = s->n_edges; info->sweep_sources = s->sweep_sources;
  info->sweep_pieces_split_rows = s->n_split_rows; info->sweep_partial_slots = s->n_partial_slots;
  info->sweep_slots_per_unit = s->slots; info->sweep_units = s->n_units; info->sweep_iterations = s->n_iters;
  info->sweep_windows = s->n_windows;
  info->rows_rows = s->plan.perm ? s->rows_rows : 0; info->rows_edges = s->plan.perm ? s->rows_edges : 0;
  info->rows_sources = s->plan.perm ? s->rows_sources : 0;
}

int launch_sweep(const lgc_graph* g, const SweepSched* s, int ld, const float* x, EpiMode mode, const EpiArgs& a,
                 float* partials, cudaStream_t st) {
  (void)g;
  if (!s || ld % 16) return LGC_ERR_INVALID;
  int rc = LGC_ERR_UNSUPPORTED;
  switch (ld / 16) {
#define LGC_SWEEP_CASE(F) case F: rc = launch_sweep_f<F>(s, x, mode, a, partials, st); break;
    LGC_SWEEP_CASE(1) LGC_SWEEP_CASE(2) LGC_SWEEP_CASE(3) LGC_SWEEP_CASE(4) LGC_SWEEP_CASE(5)
    LGC_SWEEP_CASE(6) LGC_SWEEP_CASE(8) LGC_SWEEP_CASE(10) LGC_SWEEP_CASE(12) LGC_SWEEP_CASE(16)
#undef LGC_SWEEP_CASE
    default: break;
  }
  if (rc) return rc;
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

}  // namespace lgc
