// Graph build: COO (int64, `df_to_graph` layout) -> destination-major CSR + gcn_norm weights.
// Runs once per graph; replaces the degree normalisation PyG recomputes in every LGConv call
// (reference call site src/lightgcn.py:96; input layout src/utils_v2.py:146-165).
#include <cub/cub.cuh>
#include <nvtx3/nvToolsExt.h>

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace lgc {

thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

long long g_launch_count = 0;

static bool nvtx_on() {
  static const bool on = [] { const char* e = getenv("LGC_NVTX"); return e && atoi(e) != 0; }();
  return on;
}
void nvtx_push(const char* name) { if (nvtx_on()) nvtxRangePushA(name); }
void nvtx_pop() { if (nvtx_on()) nvtxRangePop(); }

struct ProfEvent { int tag; cudaEvent_t beg, end; };
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof_events;
bool prof_enabled() { return g_prof_on; }
void prof_record(int tag, cudaStream_t st, bool begin) {
  if (begin) {
    ProfEvent e; e.tag = tag;
    cudaEventCreate(&e.beg); cudaEventCreate(&e.end);
    cudaEventRecord(e.beg, st);
    g_prof_events.push_back(e);
  } else {
    for (size_t i = g_prof_events.size(); i-- > 0;)
      if (g_prof_events[i].tag == tag) { cudaEventRecord(g_prof_events[i].end, st); break; }
  }
}

namespace {

constexpr int kLightMaxDegree = 32;  // rows up to this in-degree: one sub-warp per row
constexpr int kChunkEdges = 256;     // heavy rows are cut into warp work items of <= this

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// Where the edge list comes from. CooEdges: the `[2, nnz]` int64 `edge_index` of df_to_graph.
// PairEdges: the E interaction pairs themselves; edge e < E is a[e] -> b[e], edge E + e is
// b[e] -> a[e] with the same weight -- the order in which df_to_graph concatenates the two
// directions (src/utils_v2.py:155-158), so both sources give bit-identical graphs, without the
// [2, 2E] int64 intermediate (6.4 GB at 200 M interactions).
struct CooEdges {
  const int64_t* ei; const float* ew; int64_t nnz;
  __device__ __forceinline__ int64_t src(int64_t e) const { return ei[e]; }
  __device__ __forceinline__ int64_t dst(int64_t e) const { return ei[nnz + e]; }
  __device__ __forceinline__ float weight(int64_t e) const { return ew ? ew[e] : 1.0f; }
};
struct PairEdges {
  const int64_t* a; const int64_t* b; const float* ew; int64_t n_pairs;
  __device__ __forceinline__ int64_t src(int64_t e) const { return e < n_pairs ? a[e] : b[e - n_pairs]; }
  __device__ __forceinline__ int64_t dst(int64_t e) const { return e < n_pairs ? b[e] : a[e - n_pairs]; }
  __device__ __forceinline__ float weight(int64_t e) const { return ew ? ew[e < n_pairs ? e : e - n_pairs] : 1.0f; }
};

// keys = target id, vals = edge position; count in-degrees; range check; symmetry fingerprint
template <class Edges>
__global__ void k_extract(Edges edges, int64_t nnz,
                          int64_t num_nodes, int64_t num_cols, int32_t* __restrict__ keys, int32_t* __restrict__ vals,
                          int32_t* __restrict__ counts, int* __restrict__ bad,
                          unsigned long long* __restrict__ fp) {
  unsigned long long h_fwd = 0, h_rev = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = edges.src(e), c = edges.dst(e);
    if (r < 0 || r >= num_cols || c < 0 || c >= num_nodes) {
      atomicExch(bad, 1);
      keys[e] = 0; vals[e] = (int32_t)e;
      continue;
    }
    keys[e] = (int32_t)c;
    vals[e] = (int32_t)e;
    atomicAdd(&counts[c], 1);
    unsigned long long wb = (unsigned long long)__float_as_uint(edges.weight(e));
    h_fwd += mix64(((unsigned long long)r << 32 | (unsigned long long)c) ^ mix64(wb + 0x9e3779b97f4a7c15ULL));
    h_rev += mix64(((unsigned long long)c << 32 | (unsigned long long)r) ^ mix64(wb + 0x9e3779b97f4a7c15ULL));
  }
  // commutative 64-bit sums: equal iff the multisets {(src,dst,w)} and {(dst,src,w)} agree (w.h.p.)
  for (int o = 16; o; o >>= 1) {
    h_fwd += __shfl_xor_sync(0xffffffffu, h_fwd, o);
    h_rev += __shfl_xor_sync(0xffffffffu, h_rev, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&fp[0], h_fwd); atomicAdd(&fp[1], h_rev); }
}

template <class Edges>
__global__ void k_gather_csr(Edges edges, int64_t nnz,
                             const int32_t* __restrict__ eid, int32_t* __restrict__ src,
                             float* __restrict__ w) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nnz;
       j += (int64_t)gridDim.x * blockDim.x) {
    int32_t e = eid[j];
    src[j] = (int32_t)edges.src(e);
    w[j] = edges.weight(e);
  }
}

// Weighted in-degree, one warp per row, accumulated strictly in CSR (= edge-list) order so the
// fp32 result is bit-identical to the CPU `scatter_add_` of gcn_norm. dis = deg^-1/2, inf -> 0.
__global__ void k_degree(const int32_t* __restrict__ rowptr, const float* __restrict__ w,
                         int64_t num_nodes, float* __restrict__ deg, float* __restrict__ dis) {
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= num_nodes) return;
  int beg = rowptr[row], end = rowptr[row + 1];
  float sum = 0.0f;
  for (int base = beg; base < end; base += 32) {
    float val = (base + lane < end) ? w[base + lane] : 0.0f;
    int n = min(32, end - base);
    for (int j = 0; j < n; ++j) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, val, j));
  }
  if (lane == 0) {
    deg[row] = sum;
    float r = __fdiv_rn(1.0f, __fsqrt_rn(sum));   // == torch pow(-0.5) on CPU (sqrt then divide)
    if (isinf(r)) r = 0.0f;
    dis[row] = r;
  }
}

__global__ void k_normalise(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                            const float* __restrict__ dis, int64_t num_nodes, float* __restrict__ w) {
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= num_nodes) return;
  int beg = rowptr[row], end = rowptr[row + 1];
  float dr = dis[row];
  for (int j = beg + lane; j < end; j += 32)
    w[j] = __fmul_rn(__fmul_rn(dis[src[j]], w[j]), dr);   // (dis[src] * w) * dis[dst]
}

__global__ void k_count_chunks(const int32_t* __restrict__ rowptr, int64_t num_nodes,
                               int32_t* __restrict__ n_chunks, int32_t* __restrict__ n_slots,
                               int32_t* __restrict__ n_split) {
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row > num_nodes) return;
  int nc = 0;
  if (row < num_nodes) {
    int d = rowptr[row + 1] - rowptr[row];
    nc = d > kLightMaxDegree ? (d + kChunkEdges - 1) / kChunkEdges : 0;
  }
  n_chunks[row] = nc;
  n_slots[row] = nc > 1 ? nc : 0;
  n_split[row] = nc > 1 ? 1 : 0;
}

__global__ void k_fill_chunks(const int32_t* __restrict__ rowptr, int64_t num_nodes,
                              const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ slot_off,
                              const int32_t* __restrict__ split_off, int4* __restrict__ chunks,
                              int4* __restrict__ split_rows) {
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= num_nodes) return;
  int beg = rowptr[row], d = rowptr[row + 1] - beg;
  if (d <= kLightMaxDegree) return;
  int nc = (d + kChunkEdges - 1) / kChunkEdges;
  int per = (d + nc - 1) / nc;                    // balanced chunk length
  int c0 = chunk_off[row];
  for (int k = 0; k < nc; ++k) {
    int b = beg + k * per, e = min(beg + d, b + per);
    chunks[c0 + k] = make_int4((int)row, b, e, nc > 1 ? slot_off[row] + k : -1);
  }
  if (nc > 1) split_rows[split_off[row]] = make_int4((int)row, slot_off[row], nc, 0);
}

// segment offsets of the split (hub) rows for the segmented sort by source
__global__ void k_split_segments(const int4* __restrict__ split_rows, int n_split, const int32_t* __restrict__ rowptr,
                                 int32_t* __restrict__ seg_beg, int32_t* __restrict__ seg_end) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_split) return;
  const int row = split_rows[i].x;
  seg_beg[i] = rowptr[row];
  seg_end[i] = rowptr[row + 1];
}

// Scheduling key of a chunk: single-chunk rows first (their sources span the whole table anyway),
// then the chunks of the hub rows by their first source: chunks that run at the same time (the
// heavy kernel hands chunk c to warp c mod #warps) then read the same window of the source table,
// which is what lets L2 serve the ~3 reads of every user row.
__global__ void k_chunk_keys(const int4* __restrict__ chunks, int n_chunks, const int32_t* __restrict__ hsrc,
                             uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const int4 c = chunks[i];
  keys[i] = c.w < 0 ? 0u : (uint32_t)hsrc[c.y] + 1u;
  vals[i] = i;
}
__global__ void k_chunk_gather(const int4* __restrict__ in, const int32_t* __restrict__ order, int n,
                               int4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[order[i]];
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  T* release() { T* q = p; p = nullptr; return q; }
};

}  // namespace
}  // namespace lgc

using namespace lgc;

extern "C" int lgc_abi_version(void) { return LGC_ABI_VERSION; }
extern "C" const char* lgc_last_error(void) { return g_last_error.c_str(); }

extern "C" long long lgc_launch_count(void) { return g_launch_count; }

extern "C" int lgc_profile_enable(int on) {
  for (auto& e : g_prof_events) { cudaEventDestroy(e.beg); cudaEventDestroy(e.end); }
  g_prof_events.clear();
  g_prof_on = on != 0;
  return LGC_OK;
}

extern "C" int lgc_profile_read(double* h_ms, long long* h_count, int n_tags) {
  LGC_REQUIRE(h_ms && h_count && n_tags > 0, "bad argument");
  for (int i = 0; i < n_tags; ++i) { h_ms[i] = 0.0; h_count[i] = 0; }
  for (auto& e : g_prof_events) {
    LGC_CUDA(cudaEventSynchronize(e.end));
    float ms = 0.f;
    LGC_CUDA(cudaEventElapsedTime(&ms, e.beg, e.end));
    if (e.tag >= 0 && e.tag < n_tags) { h_ms[e.tag] += ms; h_count[e.tag] += 1; }
    cudaEventDestroy(e.beg); cudaEventDestroy(e.end);
  }
  g_prof_events.clear();
  return LGC_OK;
}

extern "C" int lgc_graph_destroy(lgc_graph_t* g) {
  if (!g) return LGC_OK;
  cudaFree(g->rowptr); cudaFree(g->src); cudaFree(g->eid); cudaFree(g->w_hat);
  cudaFree(g->deg); cudaFree(g->dis); cudaFree(g->chunks); cudaFree(g->split_rows);
  cudaFree(g->hsrc); cudaFree(g->hw);
  for (int i = 0; i < kMaxSweepScheds; ++i) lgc::sweep_destroy(g->sweep[i]);
  delete g;
  return LGC_OK;
}

extern "C" int lgc_graph_get_info(const lgc_graph_t* g, lgc_graph_info* info) {
  LGC_REQUIRE(g && info, "null argument");
  info->num_nodes = g->num_nodes;
  info->nnz = g->nnz;
  info->is_symmetric = g->is_symmetric;
  info->light_max_degree = g->light_max_degree;
  info->num_heavy_rows = g->num_heavy_rows;
  info->num_chunks = g->num_chunks;
  info->num_split_rows = g->num_split_rows;
  info->rowptr = g->rowptr; info->src = g->src; info->eid = g->eid;
  info->w_hat = g->w_hat; info->deg = g->deg; info->dis = g->dis;
  return LGC_OK;
}

template <class Edges>
static int graph_build_impl(int64_t num_nodes, int64_t num_cols, int64_t nnz, Edges edges,
                            int normalize, void* stream_, lgc_graph_t** out);

extern "C" int lgc_graph_build(int64_t num_nodes, int64_t nnz, const int64_t* ei, const float* ew,
                               int normalize, void* stream_, lgc_graph_t** out) {
  LGC_REQUIRE(nnz <= 0 || ei, "edge_index is null");
  return graph_build_impl(num_nodes, num_nodes, nnz, CooEdges{ei, ew, nnz}, normalize, stream_, out);
}

extern "C" int lgc_graph_build_rect(int64_t num_rows, int64_t num_cols, int64_t nnz, const int64_t* ei,
                                    const float* ew, void* stream_, lgc_graph_t** out) {
  LGC_REQUIRE(nnz <= 0 || ei, "edge_index is null");
  return graph_build_impl(num_rows, num_cols, nnz, CooEdges{ei, ew, nnz}, 0, stream_, out);
}

extern "C" int lgc_graph_build_pairs(int64_t num_nodes, int64_t n_pairs, const int64_t* a, const int64_t* b,
                                     const float* ew, int normalize, void* stream_, lgc_graph_t** out) {
  LGC_REQUIRE(n_pairs >= 0 && n_pairs < (1LL << 30), "n_pairs out of range");
  LGC_REQUIRE(n_pairs == 0 || (a && b), "pair arrays are null");
  return graph_build_impl(num_nodes, num_nodes, 2 * n_pairs, PairEdges{a, b, ew, n_pairs}, normalize, stream_, out);
}

// rows = targets (row 1 of edge_index) in [0, num_nodes); sources (row 0) in [0, num_cols)
template <class Edges>
static int graph_build_impl(int64_t num_nodes, int64_t num_cols, int64_t nnz, Edges edges,
                            int normalize, void* stream_, lgc_graph_t** out) {
  LGC_REQUIRE(out, "out_graph is null");
  NvtxRange nvtx("lgc_graph_build");
  *out = nullptr;
  LGC_REQUIRE(num_nodes > 0 && num_nodes < (1LL << 31) - 64, "num_nodes out of range");
  LGC_REQUIRE(num_cols > 0 && num_cols < (1LL << 31) - 64, "num_cols out of range");
  LGC_REQUIRE(!normalize || num_cols == num_nodes, "normalisation needs a square operator");
  LGC_REQUIRE(nnz >= 0 && nnz < (1LL << 31) - 64, "nnz out of range");
  cudaStream_t stream = (cudaStream_t)stream_;

  DevBuf<int32_t> keys_in, keys_out, vals_in, counts, rowptr, src, eid, n_chunks, n_slots, n_split,
      chunk_off, slot_off, split_off;
  DevBuf<float> w, deg, dis;
  DevBuf<int> bad;
  DevBuf<unsigned long long> fp;
  DevBuf<char> tmp;
  DevBuf<int4> chunks, split_rows;
  const size_t n1 = (size_t)num_nodes + 1;
  LGC_CUDA(keys_in.alloc(nnz)); LGC_CUDA(keys_out.alloc(nnz)); LGC_CUDA(vals_in.alloc(nnz));
  // src / w_hat / rowptr are padded: the light-row SpMM bulk-copies 16-byte-granular supersets
  LGC_CUDA(eid.alloc(nnz)); LGC_CUDA(src.alloc(nnz + 8)); LGC_CUDA(w.alloc(nnz + 8));
  LGC_CUDA(cudaMemsetAsync(src.p + nnz, 0, 8 * sizeof(int32_t), stream));
  LGC_CUDA(cudaMemsetAsync(w.p + nnz, 0, 8 * sizeof(float), stream));
  LGC_CUDA(counts.alloc(n1)); LGC_CUDA(rowptr.alloc(n1 + 8));
  LGC_CUDA(cudaMemsetAsync(rowptr.p + n1, 0, 8 * sizeof(int32_t), stream));
  LGC_CUDA(deg.alloc(num_nodes)); LGC_CUDA(dis.alloc(num_nodes));
  LGC_CUDA(bad.alloc(1)); LGC_CUDA(fp.alloc(2));
  LGC_CUDA(cudaMemsetAsync(counts.p, 0, n1 * sizeof(int32_t), stream));
  LGC_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  LGC_CUDA(cudaMemsetAsync(fp.p, 0, 2 * sizeof(unsigned long long), stream));

  const int threads = 256;
  const int grid_e = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(nnz, threads), 1), kNumSMs * 16);
  k_extract<<<grid_e, threads, 0, stream>>>(edges, nnz, num_nodes, num_cols, keys_in.p, vals_in.p, counts.p,
                                            bad.p, fp.p);
  LGC_LAUNCH_CHECK();

  int end_bit = 1;
  while ((1LL << end_bit) < num_nodes) ++end_bit;
  size_t tmp_sort = 0, tmp_scan = 0;
  LGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys_in.p, keys_out.p, vals_in.p, eid.p,
                                           (int)nnz, 0, end_bit, stream));
  LGC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, counts.p, rowptr.p, (int)n1, stream));
  LGC_CUDA(tmp.alloc(std::max(tmp_sort, tmp_scan)));
  size_t tmp_bytes = std::max(tmp_sort, tmp_scan);
  if (nnz > 0)   // LSD radix sort is stable: the edges of a row keep their edge-list order
    LGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, eid.p,
                                             (int)nnz, 0, end_bit, stream));
  LGC_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.p, rowptr.p, (int)n1, stream));
  k_gather_csr<<<grid_e, threads, 0, stream>>>(edges, nnz, eid.p, src.p, w.p);
  LGC_LAUNCH_CHECK();

  const int grid_rows_warp = (int)ceil_div(num_nodes * 32, threads);
  k_degree<<<grid_rows_warp, threads, 0, stream>>>(rowptr.p, w.p, num_nodes, deg.p, dis.p);
  LGC_LAUNCH_CHECK();
  if (normalize) {
    k_normalise<<<grid_rows_warp, threads, 0, stream>>>(rowptr.p, src.p, dis.p, num_nodes, w.p);
    LGC_LAUNCH_CHECK();
  }

  // heavy-row schedule
  LGC_CUDA(n_chunks.alloc(n1)); LGC_CUDA(n_slots.alloc(n1)); LGC_CUDA(n_split.alloc(n1));
  LGC_CUDA(chunk_off.alloc(n1)); LGC_CUDA(slot_off.alloc(n1)); LGC_CUDA(split_off.alloc(n1));
  const int grid_rows = (int)ceil_div((int64_t)n1, threads);
  k_count_chunks<<<grid_rows, threads, 0, stream>>>(rowptr.p, num_nodes, n_chunks.p, n_slots.p, n_split.p);
  LGC_LAUNCH_CHECK();
  LGC_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, n_chunks.p, chunk_off.p, (int)n1, stream));
  LGC_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, n_slots.p, slot_off.p, (int)n1, stream));
  LGC_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, n_split.p, split_off.p, (int)n1, stream));

  int h_bad = 0;
  int32_t h_tot[3] = {0, 0, 0};
  unsigned long long h_fp[2] = {0, 0};
  LGC_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaMemcpyAsync(&h_tot[0], chunk_off.p + num_nodes, 4, cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaMemcpyAsync(&h_tot[1], slot_off.p + num_nodes, 4, cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaMemcpyAsync(&h_tot[2], split_off.p + num_nodes, 4, cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaMemcpyAsync(h_fp, fp.p, 16, cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) {
    set_error("edge_index holds a node id outside [0, num_nodes)");   // (or a source outside [0, num_cols))
    return LGC_ERR_INDEX_RANGE;
  }
  LGC_CUDA(chunks.alloc(h_tot[0])); LGC_CUDA(split_rows.alloc(h_tot[2]));
  k_fill_chunks<<<grid_rows, threads, 0, stream>>>(rowptr.p, num_nodes, chunk_off.p, slot_off.p,
                                                   split_off.p, chunks.p, split_rows.p);
  LGC_LAUNCH_CHECK();
  // ---- heavy-row gather arrays: hub rows sorted by source; chunk list ordered by first source
  DevBuf<int32_t> hsrc, seg_beg, seg_end, ckey_vals_in, ckey_vals_out;
  DevBuf<float> hw;
  DevBuf<uint32_t> ckeys_in, ckeys_out;
  DevBuf<int4> chunks_sorted;
  DevBuf<char> tmp2;
  LGC_CUDA(hsrc.alloc(nnz + 8)); LGC_CUDA(hw.alloc(nnz + 8));
  LGC_CUDA(cudaMemcpyAsync(hsrc.p, src.p, (size_t)(nnz + 8) * 4, cudaMemcpyDeviceToDevice, stream));
  LGC_CUDA(cudaMemcpyAsync(hw.p, w.p, (size_t)(nnz + 8) * 4, cudaMemcpyDeviceToDevice, stream));
  if (h_tot[2] > 0) {
    const int ns = h_tot[2];
    LGC_CUDA(seg_beg.alloc(ns)); LGC_CUDA(seg_end.alloc(ns));
    k_split_segments<<<(int)ceil_div(ns, threads), threads, 0, stream>>>(split_rows.p, ns, rowptr.p, seg_beg.p,
                                                                         seg_end.p);
    LGC_LAUNCH_CHECK();
    size_t bytes = 0;
    LGC_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, bytes, src.p, hsrc.p, w.p, hw.p, (int)nnz, ns, seg_beg.p,
                                                 seg_end.p, stream));
    LGC_CUDA(tmp2.alloc(bytes));
    LGC_CUDA(cub::DeviceSegmentedSort::SortPairs(tmp2.p, bytes, src.p, hsrc.p, w.p, hw.p, (int)nnz, ns, seg_beg.p,
                                                 seg_end.p, stream));
  }
  if (h_tot[0] > 0) {
    const int nc = h_tot[0];
    LGC_CUDA(ckeys_in.alloc(nc)); LGC_CUDA(ckeys_out.alloc(nc));
    LGC_CUDA(ckey_vals_in.alloc(nc)); LGC_CUDA(ckey_vals_out.alloc(nc));
    LGC_CUDA(chunks_sorted.alloc(nc));
    k_chunk_keys<<<(int)ceil_div(nc, threads), threads, 0, stream>>>(chunks.p, nc, hsrc.p, ckeys_in.p,
                                                                     ckey_vals_in.p);
    LGC_LAUNCH_CHECK();
    size_t bytes = 0;
    LGC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ckeys_in.p, ckeys_out.p, ckey_vals_in.p,
                                             ckey_vals_out.p, nc, 0, 32, stream));
    DevBuf<char> tmp3;
    LGC_CUDA(tmp3.alloc(bytes));
    LGC_CUDA(cub::DeviceRadixSort::SortPairs(tmp3.p, bytes, ckeys_in.p, ckeys_out.p, ckey_vals_in.p,
                                             ckey_vals_out.p, nc, 0, 32, stream));
    k_chunk_gather<<<(int)ceil_div(nc, threads), threads, 0, stream>>>(chunks.p, ckey_vals_out.p, nc,
                                                                       chunks_sorted.p);
    LGC_LAUNCH_CHECK();
    LGC_CUDA(cudaStreamSynchronize(stream));      // tmp3 goes out of scope
    std::swap(chunks.p, chunks_sorted.p);
  }
  LGC_CUDA(cudaStreamSynchronize(stream));

  lgc_graph* g = new lgc_graph();
  g->num_nodes = num_nodes; g->nnz = nnz; g->num_cols = num_cols;
  g->is_symmetric = (num_cols == num_nodes && h_fp[0] == h_fp[1]) ? 1 : 0;
  g->light_max_degree = kLightMaxDegree;
  g->num_chunks = h_tot[0];
  g->num_partial_slots = h_tot[1];
  g->num_split_rows = h_tot[2];
  g->num_heavy_rows = h_tot[0] - h_tot[1] + h_tot[2];   // single-chunk rows + split rows
  g->rowptr = rowptr.release(); g->src = src.release(); g->eid = eid.release();
  g->w_hat = w.release(); g->deg = deg.release(); g->dis = dis.release();
  g->chunks = chunks.release(); g->split_rows = split_rows.release();
  g->hsrc = hsrc.release(); g->hw = hw.release();
  *out = g;
  return LGC_OK;
}
