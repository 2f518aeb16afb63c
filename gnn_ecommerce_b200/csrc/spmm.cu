// CSR SpMM  Y = A_hat X  with fused epilogues -- the LGConv hot kernel
// (replaces PyG gather + mul + scatter_add, K2-K5 of SURVEY.md 2.3; call site src/lightgcn.py:96).
//
// HBM-bound gather/accumulate. A table row of `ld` floats is VEC = ld/4 float4 = L lanes x V
// float4 per lane, so every gather of a neighbour row is a run of 128-bit loads that covers whole
// 32-byte sectors. Three launches per layer, all deterministic (no atomics):
//   light : rows with in-degree <= 32, natural row order, one L-lane sub-warp per row
//           (the power-law tail: ~97 % of the user rows), streaming writes stay sequential;
//   heavy : rows above that are pre-cut (graph build) into work items of <= 256 edges, one warp
//           each, 32/L edge streams per warp; rows that fit one item finish in place, split rows
//           write a partial row;
//   finish: split rows add their partials in order and apply the epilogue.
// The epilogue fuses what the reference runs as separate ATen passes: the running layer mean
// `out = out + x * alpha` (src/lightgcn.py:93,97), the backward Horner add, and dense Adam.
#include "spmm.cuh"

namespace lgc {
namespace {

// Operands of the epilogue that do not depend on the SpMM sum: loaded BEFORE the gathers so
// their latency overlaps the gather latency instead of following it.
struct Pre {
  float4 r0, r1, r2, r3;
};

template <int MODE>
__device__ __forceinline__ void epi_preload(const EpiArgs& a, size_t off, Pre& p) {
  if (MODE == EPI_PLAIN) {
    if (a.addend) p.r0 = ldg_f4(a.addend + off);
  } else if (MODE == EPI_FWD_INIT) {
    p.r0 = ldg_f4(a.xrow + off);
  } else if (MODE == EPI_FWD_RMW) {
    p.r0 = ld_f4_cs(a.acc + off);
  } else {  // EPI_ADAM
    p.r0 = ld_f4_cs(a.addend + off);
    p.r1 = ld_f4(a.p + off);
    p.r2 = ld_f4_cs(a.m + off);
    p.r3 = ld_f4_cs(a.v + off);
  }
}

template <int MODE>
__device__ __forceinline__ void epi_finish(const EpiArgs& a, size_t off, float4 s, const Pre& q) {
  if (MODE == EPI_PLAIN) {
    float4 r = make_float4(a.scale * s.x, a.scale * s.y, a.scale * s.z, a.scale * s.w);
    if (a.addend) {
      r.x = fmaf(a.beta, q.r0.x, r.x); r.y = fmaf(a.beta, q.r0.y, r.y);
      r.z = fmaf(a.beta, q.r0.z, r.z); r.w = fmaf(a.beta, q.r0.w, r.w);
    }
    st_f4(a.y + off, r);
  } else if (MODE == EPI_FWD_INIT) {
    if (a.y) st_f4(a.y + off, s);
    float4 r;                              // out = x * alpha0; out = out + x1 * alpha1
    r.x = __fadd_rn(__fmul_rn(q.r0.x, a.a0), __fmul_rn(s.x, a.a1));
    r.y = __fadd_rn(__fmul_rn(q.r0.y, a.a0), __fmul_rn(s.y, a.a1));
    r.z = __fadd_rn(__fmul_rn(q.r0.z, a.a0), __fmul_rn(s.z, a.a1));
    r.w = __fadd_rn(__fmul_rn(q.r0.w, a.a0), __fmul_rn(s.w, a.a1));
    st_f4_cs(a.acc + off, r);
  } else if (MODE == EPI_FWD_RMW) {
    if (a.y) st_f4(a.y + off, s);
    float4 o = q.r0;
    o.x = __fadd_rn(o.x, __fmul_rn(s.x, a.a1)); o.y = __fadd_rn(o.y, __fmul_rn(s.y, a.a1));
    o.z = __fadd_rn(o.z, __fmul_rn(s.z, a.a1)); o.w = __fadd_rn(o.w, __fmul_rn(s.w, a.a1));
    st_f4_cs(a.acc + off, o);
  } else {  // EPI_ADAM
    float4 p = q.r1, m = q.r2, v = q.r3;
    adam_update(p.x, m.x, v.x, fmaf(a.scale, s.x, q.r0.x), a.adam);
    adam_update(p.y, m.y, v.y, fmaf(a.scale, s.y, q.r0.y), a.adam);
    adam_update(p.z, m.z, v.z, fmaf(a.scale, s.z, q.r0.z), a.adam);
    adam_update(p.w, m.w, v.w, fmaf(a.scale, s.w, q.r0.w), a.adam);
    st_f4(a.p + off, p);
    st_f4_cs(a.m + off, m);
    st_f4_cs(a.v + off, v);
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue(const EpiArgs& a, size_t off, float4 s) {
  Pre q;
  epi_preload<MODE>(a, off, q);
  epi_finish<MODE>(a, off, s, q);
}

// ---------------------------------------------------------------------------------- light rows
// One CTA = a tile of consecutive rows. The tile's rowptr slice and (when it fits) its contiguous
// CSR slice of (source, weight) pairs are staged in shared memory with two coalesced rounds, so a
// row costs ONE exposed memory latency (its gathers, issued together with its epilogue operands)
// instead of four dependent ones (rowptr -> indices -> gathers -> epilogue operands). Each L-lane
// sub-warp walks its rows two at a time to double the loads in flight.
constexpr int kLightRowsPerSub = 8;
constexpr int kLightStageCap = 2048;   // staged CSR entries per tile (16 KB)

template <int L, int V, int MODE>
__global__ void __launch_bounds__(256) k_spmm_light(const int32_t* __restrict__ rowptr,
                                                    const int32_t* __restrict__ src,
                                                    const float* __restrict__ w,
                                                    const float* __restrict__ x, int num_rows,
                                                    int light_max, EpiArgs args) {
  constexpr int LD = 4 * L * V;
  constexpr int NSUB = 256 / L;                       // sub-warps per CTA
  constexpr int TILE = NSUB * kLightRowsPerSub;       // rows per CTA
  __shared__ int s_rp[TILE + 1];
  __shared__ int s_src[kLightStageCap];
  __shared__ float s_w[kLightStageCap];

  const int row0 = blockIdx.x * TILE;
  for (int t = threadIdx.x; t <= TILE; t += 256) s_rp[t] = rowptr[min(row0 + t, num_rows)];
  __syncthreads();
  const int e0 = s_rp[0], n_e = s_rp[TILE] - e0;
  const bool staged = n_e <= kLightStageCap;
  if (staged) {
    for (int i = threadIdx.x; i < n_e; i += 256) { s_src[i] = src[e0 + i]; s_w[i] = w[e0 + i]; }
  }
  __syncthreads();

  const int sw = threadIdx.x / L, sl = threadIdx.x % L;
#pragma unroll 1
  for (int j = 0; j < kLightRowsPerSub; j += 2) {
    // consecutive sub-warps take consecutive rows: a warp's stores are contiguous
    const int ta = j * NSUB + sw, tb = ta + NSUB;
    const int rowa = row0 + ta, rowb = row0 + tb;
    int bega = s_rp[ta], dega = s_rp[ta + 1] - bega;
    int begb = s_rp[tb], degb = s_rp[tb + 1] - begb;
    const bool oka = rowa < num_rows && dega <= light_max;
    const bool okb = rowb < num_rows && degb <= light_max;
    if (!oka) dega = 0;
    if (!okb) degb = 0;
    const size_t offa = (size_t)rowa * LD + 4 * sl, offb = (size_t)rowb * LD + 4 * sl;
    Pre qa[V], qb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (oka) epi_preload<MODE>(args, offa + 4 * L * v, qa[v]);
      if (okb) epi_preload<MODE>(args, offb + 4 * L * v, qb[v]);
    }
    float4 acca[V], accb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acca[v] = accb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int dmax = max(dega, degb);
    for (int t = 0; t < dmax; t += 2) {
      int sa[2], sb[2]; float wa[2], wb[2]; float4 xa[2][V], xb[2][V];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (t + u < dega) {
          const int e = bega + t + u;
          sa[u] = staged ? s_src[e - e0] : src[e];
          wa[u] = staged ? s_w[e - e0] : w[e];
        }
        if (t + u < degb) {
          const int e = begb + t + u;
          sb[u] = staged ? s_src[e - e0] : src[e];
          wb[u] = staged ? s_w[e - e0] : w[e];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (t + u < dega) {
          const float* xr = x + (size_t)sa[u] * LD + 4 * sl;
#pragma unroll
          for (int v = 0; v < V; ++v) xa[u][v] = ldg_f4(xr + 4 * L * v);
        }
        if (t + u < degb) {
          const float* xr = x + (size_t)sb[u] * LD + 4 * sl;
#pragma unroll
          for (int v = 0; v < V; ++v) xb[u][v] = ldg_f4(xr + 4 * L * v);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (t + u < dega) {
#pragma unroll
          for (int v = 0; v < V; ++v) acca[v] = fma4(wa[u], xa[u][v], acca[v]);
        }
        if (t + u < degb) {
#pragma unroll
          for (int v = 0; v < V; ++v) accb[v] = fma4(wb[u], xb[u][v], accb[v]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (oka) epi_finish<MODE>(args, offa + 4 * L * v, acca[v], qa[v]);
      if (okb) epi_finish<MODE>(args, offb + 4 * L * v, accb[v], qb[v]);
    }
  }
}

// ---------------------------------------------------------------------------------- heavy rows
template <int L, int V, int MODE>
__global__ void __launch_bounds__(256, (V == 1) ? 3 : 2)
k_spmm_heavy(const int4* __restrict__ chunks, int num_chunks, const int32_t* __restrict__ src,
             const float* __restrict__ w, const float* __restrict__ x, float* __restrict__ partials,
             EpiArgs args) {
  constexpr int RPW = 32 / L;        // edge streams per warp
  constexpr int LD = 4 * L * V;
  constexpr int U = (L >= 8) ? 8 : L;  // gathers in flight per stream and unrolled step
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, sl = lane % L;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (warp >= num_chunks) return;
  const int4 c = chunks[warp];
  const int row = c.x, beg = c.y, end = c.z, slot = c.w;

  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = beg; base < end; base += 32) {
    int my_s = 0; float my_w = 0.f;
    if (base + lane < end) { my_s = src[base + lane]; my_w = w[base + lane]; }
    const int n = min(32, end - base);
    // stream `sub` takes entries sub, sub+RPW, ... of this block of 32
#pragma unroll
    for (int t0 = 0; t0 < L; t0 += U) {
      float ww[U]; float4 xv[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = (t0 + u) * RPW + sub;
        const int s = __shfl_sync(0xffffffffu, my_s, idx);
        ww[u] = __shfl_sync(0xffffffffu, my_w, idx);
        if (idx < n) {
          const float* xr = x + (size_t)s * LD + 4 * sl;
#pragma unroll
          for (int v = 0; v < V; ++v) xv[u][v] = ldg_f4(xr + 4 * L * v);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if ((t0 + u) * RPW + sub < n) {
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = fma4(ww[u], xv[u][v], acc[v]);
        }
    }
  }
  // combine the RPW edge streams (fixed order: deterministic)
#pragma unroll
  for (int o = L; o < 32; o <<= 1) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
      acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
      acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
      acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
    }
  }
  if (sub != 0) return;
  if (slot >= 0) {
    float* pr = partials + (size_t)slot * LD + 4 * sl;
#pragma unroll
    for (int v = 0; v < V; ++v) st_f4(pr + 4 * L * v, acc[v]);
  } else {
    const size_t off = (size_t)row * LD + 4 * sl;
#pragma unroll
    for (int v = 0; v < V; ++v) epilogue<MODE>(args, off + 4 * L * v, acc[v]);
  }
}

// One CTA per split row: 256/L streams add the row's partials (fixed assignment and a fixed
// shared-memory reduction order: deterministic), then sub-warp 0 applies the epilogue.
template <int L, int V, int MODE>
__global__ void __launch_bounds__(256) k_spmm_finish(const int4* __restrict__ split_rows, int num_split,
                                                     const float* __restrict__ partials, EpiArgs args) {
  constexpr int LD = 4 * L * V;
  constexpr int NSUB = 256 / L;
  __shared__ float4 s_red[NSUB][L * V];
  const int4 r = split_rows[blockIdx.x];
  const int sw = threadIdx.x / L, sl = threadIdx.x % L;
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k = sw; k < r.z; k += NSUB) {
    const float* pr = partials + (size_t)(r.y + k) * LD + 4 * sl;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], ld_f4(pr + 4 * L * v));
  }
#pragma unroll
  for (int v = 0; v < V; ++v) s_red[sw][sl + L * v] = acc[v];
  __syncthreads();
  if (sw != 0) return;
  const int used = min(NSUB, r.z);
  for (int k = 1; k < used; ++k) {
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], s_red[k][sl + L * v]);
  }
  const size_t off = (size_t)r.x * LD + 4 * sl;
#pragma unroll
  for (int v = 0; v < V; ++v) epilogue<MODE>(args, off + 4 * L * v, acc[v]);
}

template <int L, int V, int MODE>
int launch_lv(const lgc_graph* g, const float* x, const EpiArgs& a, float* partials, cudaStream_t st) {
  const int threads = 256, wpb = threads / 32;
  const int64_t n = g->num_nodes;
  const int grid_light = (int)ceil_div(n, (256 / L) * kLightRowsPerSub);
  {
    ProfScope ps(PROF_LIGHT + MODE, st);
    k_spmm_light<L, V, MODE><<<grid_light, threads, 0, st>>>(g->rowptr, g->src, g->w_hat, x, (int)n,
                                                              g->light_max_degree, a);
  }
  LGC_LAUNCH_CHECK();
  if (g->num_chunks > 0) {
    const int grid_heavy = (int)ceil_div(g->num_chunks, wpb);
    {
      ProfScope ps(PROF_HEAVY + MODE, st);
      k_spmm_heavy<L, V, MODE><<<grid_heavy, threads, 0, st>>>(g->chunks, (int)g->num_chunks, g->src,
                                                                g->w_hat, x, partials, a);
    }
    LGC_LAUNCH_CHECK();
  }
  if (g->num_split_rows > 0) {
    const int grid_fin = (int)g->num_split_rows;
    {
      ProfScope ps(PROF_FINISH + MODE, st);
      k_spmm_finish<L, V, MODE><<<grid_fin, threads, 0, st>>>(g->split_rows, (int)g->num_split_rows,
                                                               partials, a);
    }
    LGC_LAUNCH_CHECK();
  }
  return LGC_OK;
}

template <int L, int V>
int launch_mode(const lgc_graph* g, const float* x, EpiMode mode, const EpiArgs& a, float* partials,
                cudaStream_t st) {
  switch (mode) {
    case EPI_PLAIN: return launch_lv<L, V, EPI_PLAIN>(g, x, a, partials, st);
    case EPI_FWD_INIT: return launch_lv<L, V, EPI_FWD_INIT>(g, x, a, partials, st);
    case EPI_FWD_RMW: return launch_lv<L, V, EPI_FWD_RMW>(g, x, a, partials, st);
    case EPI_ADAM: return launch_lv<L, V, EPI_ADAM>(g, x, a, partials, st);
  }
  return LGC_ERR_INVALID;
}

__global__ void k_scale(const float4* __restrict__ x, float4* __restrict__ y, float a, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    y[i] = make_float4(v.x * a, v.y * a, v.z * a, v.w * a);
  }
}

}  // namespace

size_t spmm_partials_floats(const lgc_graph* g, int ld) {
  return (size_t)g->num_partial_slots * (size_t)ld;
}

int launch_spmm(const lgc_graph* g, int ld, const float* x, EpiMode mode, const EpiArgs& a,
                float* partials, cudaStream_t st) {
  RowShape rs;
  if (!row_shape(ld, &rs)) {
    set_error("unsupported row width ld=" + std::to_string(ld));
    return LGC_ERR_UNSUPPORTED;
  }
#define LGC_CASE(LL, VV) \
  if (rs.L == LL && rs.V == VV) return launch_mode<LL, VV>(g, x, mode, a, partials, st);
  LGC_CASE(16, 1) LGC_CASE(16, 2) LGC_CASE(16, 3) LGC_CASE(16, 4)
  LGC_CASE(8, 1) LGC_CASE(8, 3) LGC_CASE(8, 5)
  LGC_CASE(4, 1) LGC_CASE(4, 3) LGC_CASE(4, 5)
  LGC_CASE(2, 1) LGC_CASE(1, 1)
#undef LGC_CASE
  set_error("no kernel instantiation for ld=" + std::to_string(ld));
  return LGC_ERR_UNSUPPORTED;
}

}  // namespace lgc

using namespace lgc;

extern "C" int lgc_ld_supported(int ld) {
  RowShape rs;
  if (!row_shape(ld, &rs)) return 0;
  const int ok[][2] = {{16, 1}, {16, 2}, {16, 3}, {16, 4}, {8, 1}, {8, 3}, {8, 5},
                       {4, 1},  {4, 3},  {4, 5},  {2, 1},  {1, 1}};
  for (auto& p : ok)
    if (p[0] == rs.L && p[1] == rs.V) return 1;
  return 0;
}

extern "C" size_t lgc_spmm_workspace_bytes(const lgc_graph_t* g, int ld) {
  return g ? spmm_partials_floats(g, ld) * sizeof(float) : 0;
}

extern "C" int lgc_spmm(const lgc_graph_t* g, int ld, const float* x, float* y, void* workspace,
                        size_t workspace_bytes, void* stream) {
  LGC_REQUIRE(g && x && y, "null argument");
  LGC_REQUIRE(x != y, "x and y must not alias");
  if (workspace_bytes < lgc_spmm_workspace_bytes(g, ld)) {
    set_error("lgc_spmm: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  EpiArgs a;
  a.y = y;
  return launch_spmm(g, ld, x, EPI_PLAIN, a, (float*)workspace, (cudaStream_t)stream);
}

extern "C" size_t lgc_propagate_workspace_bytes(const lgc_graph_t* g, int ld, int num_layers) {
  if (!g) return 0;
  size_t t = (size_t)g->num_nodes * ld;
  size_t n_tmp = num_layers > 2 ? 2 : (num_layers > 1 ? 1 : 0);
  return (n_tmp * t + spmm_partials_floats(g, ld)) * sizeof(float);
}

extern "C" int lgc_propagate(const lgc_graph_t* g, int ld, int num_layers, const float* h_alpha,
                             const float* x0, float* out, void* workspace, size_t workspace_bytes,
                             void* stream) {
  LGC_REQUIRE(g && h_alpha && x0 && out, "null argument");
  LGC_REQUIRE(num_layers >= 0, "num_layers < 0");
  LGC_REQUIRE(x0 != out, "x0 and out must not alias");
  if (workspace_bytes < lgc_propagate_workspace_bytes(g, ld, num_layers)) {
    set_error("lgc_propagate: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t t = (size_t)g->num_nodes * ld;
  if (num_layers == 0) {
    k_scale<<<kNumSMs * 8, 256, 0, st>>>((const float4*)x0, (float4*)out, h_alpha[0], (int64_t)(t / 4));
    LGC_LAUNCH_CHECK();
    return LGC_OK;
  }
  float* ws = (float*)workspace;
  float* tmp[2] = {ws, ws + t};
  float* partials = ws + (num_layers > 2 ? 2 : (num_layers > 1 ? 1 : 0)) * t;
  const float* cur = x0;
  for (int l = 1; l <= num_layers; ++l) {
    EpiArgs a;
    a.acc = out;
    a.a1 = h_alpha[l];
    a.y = (l < num_layers) ? tmp[(l - 1) & 1] : nullptr;   // the last layer's x is never read
    int rc;
    if (l == 1) {
      a.a0 = h_alpha[0];
      a.xrow = x0;
      rc = launch_spmm(g, ld, cur, EPI_FWD_INIT, a, partials, st);
    } else {
      rc = launch_spmm(g, ld, cur, EPI_FWD_RMW, a, partials, st);
    }
    if (rc) return rc;
    cur = a.y;
  }
  return LGC_OK;
}
