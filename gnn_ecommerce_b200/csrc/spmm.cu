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
#include <algorithm>
#include <cstdlib>
#include <type_traits>

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
    const AdamScalars ad = a.adam_dev ? *a.adam_dev : a.adam;
    const float ib = __frcp_rn(ad.bc2_sqrt);
    adam_update_fast(p.x, m.x, v.x, fmaf(a.scale, s.x, q.r0.x), ad, ib);
    adam_update_fast(p.y, m.y, v.y, fmaf(a.scale, s.y, q.r0.y), ad, ib);
    adam_update_fast(p.z, m.z, v.z, fmaf(a.scale, s.z, q.r0.z), ad, ib);
    adam_update_fast(p.w, m.w, v.w, fmaf(a.scale, s.w, q.r0.w), ad, ib);
    st_f4(a.p + off, p);
    st_f4_cs(a.m + off, m);
    st_f4_cs(a.v + off, v);
  }
}

__device__ __forceinline__ float4 final_mean(const float4* q, const EpiArgs& a, float4 s) {
  float4 r = make_float4(__fmul_rn(q[0].x, a.ah[0]), __fmul_rn(q[0].y, a.ah[0]), __fmul_rn(q[0].z, a.ah[0]),
                         __fmul_rn(q[0].w, a.ah[0]));
#pragma unroll
  for (int i = 1; i < kMaxHist; ++i)
    if (i < a.n_hist) {
      r.x = __fadd_rn(r.x, __fmul_rn(q[i].x, a.ah[i])); r.y = __fadd_rn(r.y, __fmul_rn(q[i].y, a.ah[i]));
      r.z = __fadd_rn(r.z, __fmul_rn(q[i].z, a.ah[i])); r.w = __fadd_rn(r.w, __fmul_rn(q[i].w, a.ah[i]));
    }
  r.x = __fadd_rn(r.x, __fmul_rn(s.x, a.a1)); r.y = __fadd_rn(r.y, __fmul_rn(s.y, a.a1));
  r.z = __fadd_rn(r.z, __fmul_rn(s.z, a.a1)); r.w = __fadd_rn(r.w, __fmul_rn(s.w, a.a1));
  return r;
}

template <int MODE>
__device__ __forceinline__ void epilogue(const EpiArgs& a, size_t off, float4 s) {
  if (MODE == EPI_FWD_FINAL) {
    float4 q[kMaxHist];
#pragma unroll
    for (int i = 0; i < kMaxHist; ++i)
      if (i < a.n_hist) q[i] = ldg_f4(a.hist[i] + off);
    st_f4_cs(a.acc + off, final_mean(q, a, s));
    return;
  }
  Pre q;
  epi_preload<MODE>(a, off, q);
  epi_finish<MODE>(a, off, s, q);
}

// ---------------------------------------------------------------------------------- light rows
// Persistent, warp-autonomous: no CTA barrier anywhere. Every warp walks its own sequence of row
// tiles (TR consecutive rows, ~2 KB per operand). The epilogue operands of a tile (x / acc /
// addend / p, m, v rows) are CONTIGUOUS in HBM, so threads never load them: lane 0 issues
// bulk-async copies (cp.async.bulk, the TMA engine) of whole tiles into the warp's two-stage
// shared-memory ring, two tiles ahead, plus the tile's rowptr slice and its (src, w) slice. The
// gathers are edge-balanced (see the kernel body): the tile's edges are compacted one lane per edge
// and split evenly over the sub-warps, products accumulate in a shared-memory sum tile, the
// epilogue runs in shared memory in place, and lane 0 sends the result tiles back with bulk-async
// stores. Measured (profiles/r1f_light_bottleneck.md): the kernel is bound by issued instructions,
// not by the gather loads -- removing every gather load changes its time by 3 %.
// Rows above `light_max` (hub rows) belong to the heavy-row kernels: this kernel neither sums nor
// writes them (a tile that contains hub rows leaves as one bulk store per run of light rows).
constexpr int kLightWarpsMax = 5;      // warps per CTA (each fully independent): see LightCfg::WARPS
constexpr size_t kMaxDynamicSmem = 227 * 1024;   // per CTA on sm_100
constexpr int kLightStageCap = 128;    // CSR entries of a tile staged per warp and stage

template <int L, int V, int MODE>
struct LightCfg {
  static constexpr int LD = 4 * L * V;
  static constexpr int NSUBW = 32 / L;                                  // sub-warps per warp
  static constexpr int RPS0 = (MODE == EPI_ADAM ? 256 : 512) / LD / NSUBW;
  static constexpr int RPS = RPS0 >= 8 ? 8 : (RPS0 >= 4 ? 4 : 2);      // rows per sub-warp and tile
  static constexpr int TR = RPS * NSUBW;                                // rows per tile (multiple of 4)
  static constexpr int NBUF = MODE == EPI_ADAM ? 4 : (MODE == EPI_PLAIN ? 1 : 2);   // FWD_FINAL: n_hist (run time)
  static constexpr int TILE_FLOATS = TR * LD;
  static constexpr int RP_INTS = TR + 4;                                // rowptr slice, 16-byte granular
  static constexpr int CSR_INTS = kLightStageCap + 4;                   // aligned superset of the slice
  // per stage: operand tiles | rowptr slice | src slice | w slice
  static constexpr size_t stage_bytes(int nbuf) {
    return (size_t)nbuf * TILE_FLOATS * 4 + RP_INTS * 4 + 2 * CSR_INTS * 4;
  }
  // after the ring: 6 mbarriers | row sums S [TR x LD] | carry rows C [NSUBW x LD] | sum slot of every flat
  // edge (uint16) | s_P, s_D | carry row ids | phase clocks
  static constexpr size_t ROWID_BYTES = (size_t)kLightStageCap * 2;  // uint16 per flat edge (or padding slot) of a pass
  static constexpr size_t SCRATCH_BYTES =
      (size_t)TILE_FLOATS * 4 + (size_t)NSUBW * LD * 4 + ROWID_BYTES + 2 * 33 * 4 + 8 + 32 + 64;
  static constexpr size_t warp_bytes(int nbuf) { return 2 * stage_bytes(nbuf) + 64 + SCRATCH_BYTES; }
  // ADAM tiles carry 4 operand streams: 2 CTAs x 5 warps fill the 227 KB, 3 CTAs x 4 warps do not fit
  static constexpr int WARPS = MODE == EPI_ADAM ? kLightWarpsMax : 4;   // wanted; the launcher lowers it until the CTA fits
  static constexpr size_t smem(int nbuf, int warps) { return warps * warp_bytes(nbuf) + 128; }
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Tile streams are touched once per launch: evict-first in L2, so they do not push the gather
// sources (the item table, L2-resident) out.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar,
                                          uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(gdst), "r"(smem_addr(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Diagnostics (LGC_LIGHT_PHASES=1): device counters of the light kernel's per-phase cycles.
unsigned long long* light_phase_buffer() {
  static unsigned long long* buf = [] {
    unsigned long long* p = nullptr;
    const char* e = getenv("LGC_LIGHT_PHASES");
    if (e && atoi(e) && cudaMalloc(&p, 8 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(p, 0, 8 * sizeof(unsigned long long));
    return p;
  }();
  return buf;
}

// Requires rowptr / src / w allocations padded by >= 8 elements (graph build does that): the
// bulk copies move 16-byte-granular supersets of the slices they need.
template <int L, int V, int MODE>
__global__ void __launch_bounds__(32 * kLightWarpsMax)
k_spmm_light(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const float* __restrict__ w,
             const float* __restrict__ x, int num_rows, int light_max, unsigned long long* phases,
             EpiArgs args) {
  using C = LightCfg<L, V, MODE>;
  constexpr int LD = C::LD, NSUBW = C::NSUBW, RPS = C::RPS, TR = C::TR, TF = C::TILE_FLOATS;
  const int NBUF = MODE == EPI_FWD_FINAL ? args.n_hist : C::NBUF;
  const size_t stage_bytes = C::stage_bytes(NBUF);
  extern __shared__ __align__(128) uint8_t smem_light[];
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // plain offsets from the (128-byte aligned) dynamic shared-memory base: integer round-ups here
  // make the compiler lose the address space and emit generic LD/ST with 64-bit addresses
  uint8_t* wbase = smem_light + (size_t)wic * C::warp_bytes(NBUF);
  auto stage_tiles = [&](int st) { return reinterpret_cast<float*>(wbase + (size_t)st * stage_bytes); };
  auto stage_rp = [&](int st) { return reinterpret_cast<int*>(stage_tiles(st) + NBUF * TF); };
  auto stage_src = [&](int st) { return stage_rp(st) + C::RP_INTS; };
  auto stage_w = [&](int st) { return reinterpret_cast<float*>(stage_src(st) + C::CSR_INTS); };
  // barriers: [0..1] operand tiles, [2..3] rowptr slice, [4..5] CSR slice (index + stage)
  const uint32_t bar0 = smem_addr(wbase + 2 * stage_bytes);
  float* const s_sum = reinterpret_cast<float*>(wbase + 2 * stage_bytes + 64);
  float* const s_carry = s_sum + TF;
  uint16_t* const c_dst = reinterpret_cast<uint16_t*>(s_carry + NSUBW * LD);   // sum slot of every flat edge
  int* const s_P = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(c_dst) + C::ROWID_BYTES);   // [33] flat-edge prefix (s_P[TR] unused)
  int* const s_D = s_P + 33;                                            // [33] CSR position - flat position
  int* const s_carry_row = s_D + 35;
  // -DLGC_PHASE_CLOCKS (diagnostic build): lane 0 sums clock64() deltas per phase of the tile loop
#ifdef LGC_PHASE_CLOCKS
  unsigned long long* const s_phase = reinterpret_cast<unsigned long long*>(s_carry_row + 8);
  long long t_prev = 0;
  if (phases && lane == 0) {
    for (int i = 0; i < 8; ++i) s_phase[i] = 0;
    t_prev = clock64();
  }
#define LGC_PHASE(i)                                   \
  if (phases && lane == 0) {                           \
    const long long t_now = clock64();                 \
    s_phase[i] += (unsigned long long)(t_now - t_prev); \
    t_prev = t_now;                                    \
  }
#else
#define LGC_PHASE(i)
#endif
  const uint64_t pol = policy_evict_first();

  const int n_tiles = (num_rows + TR - 1) / TR;
  const int wpc = blockDim.x >> 5;
  const int gw = blockIdx.x * wpc + wic, nw = gridDim.x * wpc;
  const float* in0 = MODE == EPI_PLAIN ? args.addend : (MODE == EPI_FWD_INIT ? args.xrow
                    : (MODE == EPI_FWD_RMW ? args.acc : (MODE == EPI_FWD_FINAL ? args.hist[0] : args.addend)));
  const uint32_t n_in = MODE == EPI_FWD_FINAL ? (uint32_t)args.n_hist : (in0 ? 1u : 0u) + (MODE == EPI_ADAM ? 3u : 0u);

  // lane 0: operand tiles + rowptr slice of `tile` -> ring stage `st` (two tiles ahead)
  auto request_tile = [&](int tile, int st) {
    if (tile >= n_tiles) return;
    const int row0 = tile * TR;
    const int rows = min(TR, num_rows - row0);
    const uint32_t rp_bytes = (uint32_t)((rows + 1 + 3) & ~3) * 4;
    mbar_expect(bar0 + 8u * (2 + st), rp_bytes);
    bulk_load(stage_rp(st), rowptr + row0, rp_bytes, bar0 + 8u * (2 + st), pol);
    if (n_in == 0) return;
    const uint32_t bytes = (uint32_t)rows * LD * 4;
    const size_t goff = (size_t)row0 * LD;
    float* b = stage_tiles(st);
    const uint32_t bar = bar0 + 8u * st;
    mbar_expect(bar, n_in * bytes);
    if (in0) bulk_load(b, in0 + goff, bytes, bar, pol);
    if (MODE == EPI_ADAM) {
      bulk_load(b + TF, args.p + goff, bytes, bar, pol);
      bulk_load(b + 2 * TF, args.m + goff, bytes, bar, pol);
      bulk_load(b + 3 * TF, args.v + goff, bytes, bar, pol);
    }
    if (MODE == EPI_FWD_FINAL) {
      for (int i = 1; i < args.n_hist; ++i) bulk_load(b + i * TF, args.hist[i] + goff, bytes, bar, pol);
    }
  };
  // lane 0: once the rowptr slice of `tile` (use count `k` of stage `st`) has landed, request the
  // 16-byte-aligned superset of its (src, w) slice (one tile ahead)
  auto request_csr = [&](int tile, int st, int k) {
    if (tile >= n_tiles) return;
    mbar_wait_parity(bar0 + 8u * (2 + st), (uint32_t)(k >> 1) & 1u);
    const int rows = min(TR, num_rows - tile * TR);
    const int* rp = stage_rp(st);
    const int e0 = rp[0], e1 = rp[rows];
    const int a0 = e0 & ~3, a1 = (e1 + 3) & ~3;
    if (a1 - a0 > C::CSR_INTS || a1 == a0) return;          // too long (read from global) or empty
    const uint32_t bytes = (uint32_t)(a1 - a0) * 4;
    mbar_expect(bar0 + 8u * (4 + st), 2 * bytes);
    bulk_load(stage_src(st), src + a0, bytes, bar0 + 8u * (4 + st), pol);
    bulk_load(stage_w(st), w + a0, bytes, bar0 + 8u * (4 + st), pol);
  };
  if (lane == 0) {
    for (int b = 0; b < 6; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    request_tile(gw, 0);
    request_tile(gw + nw, 1);
    request_csr(gw, 0, 0);
  }
  __syncwarp();

  const int sw = lane / L, sl = lane % L;
  int k = 0;
  uint32_t csr_par = 0;                                  // phase parity of the two CSR barriers
  for (int tile = gw; tile < n_tiles; tile += nw, ++k) {
    const int st = k & 1;
    const uint32_t par = (uint32_t)(k >> 1) & 1u;
    const int row0 = tile * TR;
    const int rows_here = min(TR, num_rows - row0);
    float* buf0 = stage_tiles(st);
    float* buf1 = buf0 + (NBUF > 1 ? TF : 0);
    float* buf2 = buf0 + (NBUF > 2 ? 2 * TF : 0);
    float* buf3 = buf0 + (NBUF > 3 ? 3 * TF : 0);
    const int* s_rp = stage_rp(st);

    // ---- the tile's rowptr slice and (when it fits) its CSR slice: already in shared memory
    mbar_wait_parity(bar0 + 8u * (2 + st), par);
    const int e0 = s_rp[0], n_e = s_rp[rows_here] - e0;
    const int a0 = e0 & ~3, a1 = (e0 + n_e + 3) & ~3;   // staged entries start at the aligned edge
    const bool staged = (a1 - a0 <= C::CSR_INTS) && (a1 != a0);   // == "request_csr issued a copy"
    if (staged) {                                        // this barrier only advances when used
      mbar_wait_parity(bar0 + 8u * (4 + st), (csr_par >> st) & 1u);
      csr_par ^= 1u << st;
    }
    LGC_PHASE(0)

    // ---- gathers, EDGE-balanced. One sub-warp per row made the warp wait for the longest of the
    // tile's rows (power law) with most lanes idle, and the kernel is bound by issued instructions.
    // Instead the tile's light edges form one flat list (exclusive prefix s_P over the rows; rows
    // above light_max contribute nothing), handled in passes of <= CAP edges:
    //  1. compaction, one LANE per edge: row by bisection, then (source row offset, weight,
    //     shared-memory slot of the sum) are written over the staged CSR slice in flat order;
    //  2. the pass is cut into NSUBW equal runs (padded to a multiple of U with weight-0 slots), one
    //     per sub-warp, U gathers in flight each; the sum of the current slot stays in registers and
    //     is stored when the run leaves the slot. Every slot has ONE writer per pass: a row that began
    //     in an earlier run of the pass goes to the sub-warp's carry slot instead, added afterwards in
    //     sub-warp order (timing-independent), so slots start from zero and need no clearing.
    int n_flat;
    unsigned okbits, zbits;
    {
      int dgr = 0;
      if (lane < rows_here) dgr = s_rp[lane + 1] - s_rp[lane];
      const bool light = lane < rows_here && dgr <= light_max;
      okbits = __ballot_sync(0xffffffffu, light);
      zbits = __ballot_sync(0xffffffffu, light && dgr == 0);   // rows without edges: their sum slot is never written
      const int dl = light ? dgr : 0;
      int incl = dl;
#pragma unroll
      for (int o = 1; o < TR; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      n_flat = __shfl_sync(0xffffffffu, incl, TR - 1);
      if (lane < TR) {
        s_P[lane] = incl - dl;
        s_D[lane] = (lane < rows_here ? s_rp[lane] : 0) - (incl - dl);   // CSR position = flat position + s_D[row]
      }
    }
    {
      constexpr int U0 = V >= 3 ? 2 : 4;                   // gathers in flight per lane: U * V float4
      constexpr int U = NSUBW * U0 > 32 ? 32 / NSUBW : U0;
      constexpr int CAP = kLightStageCap - 32;             // flat edges per pass; a multiple of NSUBW * U
      static_assert(CAP % (NSUBW * U) == 0, "full passes need no padding");
      int* const c_off = stage_src(st);                    // compacted in place: source row offset / 16 B
      float* const c_w = stage_w(st);
      const char* xb = reinterpret_cast<const char*>(x) + 16 * sl;
      asm("" : "+l"(xb));                                  // one 64-bit register pair: IMAD.WIDE adds it directly
      char* const sum_b = reinterpret_cast<char*>(s_sum) + 16 * sl;
      for (int base = 0; base < n_flat; base += CAP) {
        const int cnt = min(CAP, n_flat - base);
        // run length per sub-warp, a multiple of U: the slots past the pass's last edge are padding
        // (weight 0, source row 0, sum slot = the unused carry slot of sub-warp 0)
        const int len = ((cnt + NSUBW - 1) / NSUBW + U - 1) / U * U;
        const int n_slots = NSUBW * len;
        if (lane < NSUBW) s_carry_row[lane] = -1;
        __syncwarp();
        for (int j0 = 0; j0 < n_slots; j0 += 32) {
          const int j = j0 + lane;
          int off16 = 0, dst = TF * 4;
          float wgt = 0.f;
          if (j < cnt) {
            const int f = base + j;
            int r = 0;
#pragma unroll
            for (int step = TR / 2; step >= 1; step >>= 1)
              if (s_P[r + step] <= f) r += step;
            const int e = f + s_D[r];
            const int sidx = staged ? c_off[e - a0] : src[e];
            wgt = staged ? c_w[e - a0] : w[e];
            off16 = sidx * (LD / 4);
            int q = 0;
#pragma unroll
            for (int k = 1; k < NSUBW; ++k) q += (j >= k * len) ? 1 : 0;
            const bool carry = q > 0 && s_P[r] < base + q * len;   // the row began in an earlier run of this pass
            dst = (carry ? TF + q * LD : r * LD) * 4;
            if (base > 0 && q == 0 && s_P[r] < base) dst |= 1;   // continued from the previous pass: keep its sum
            if (carry && j == q * len) s_carry_row[q] = r;
          }
          __syncwarp();                                    // in place: all reads of the batch before its writes
          if (j < n_slots) {
            c_off[j] = off16;
            c_w[j] = wgt;
            c_dst[j] = (uint16_t)dst;
          }
        }
        __syncwarp();
        const int n_steps = len / U;                       // same for every sub-warp
        // The running sum of the current slot lives in registers: shared memory is read when the
        // slot changes and written when it is left (the kernel is bound by shared-memory wavefronts:
        // a read-modify-write per edge cost 16 of them, rows have ~3 edges).
        int cur = -1;
        float4 acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0, j = sw * len; it < n_steps; ++it, j += U) {
          int dsto[U]; float wv[U]; float4 xv[U][V];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float* xr = reinterpret_cast<const float*>(xb + (size_t)(unsigned)c_off[j + u] * 16);
            wv[u] = c_w[j + u];
            dsto[u] = c_dst[j + u];
#pragma unroll
            for (int v = 0; v < V; ++v) xv[u][v] = ldg_f4_ptx(xr + 4 * L * v);
          }
          // Scheduling fence. ptxas sinks each gather next to its use (one exposed round trip per
          // edge) and drops warp barriers in converged code, so the first slot offset is made to
          // depend on the last word of every gather: min() with a value >= 0xffff0000 never changes
          // the 16-bit offset, but the loads must all be in flight before the first use.
          {
            unsigned dep = 0;
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int v = 0; v < V; ++v) dep |= __float_as_uint(xv[u][v].w);
            asm("{\n.reg .u32 t;\nor.b32 t, %1, 0xffff0000;\nmin.u32 %0, %0, t;\n}" : "+r"(dsto[0]) : "r"(dep));   // opaque to the optimiser
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {                    // edge order kept per row
            if (dsto[u] != cur) {
              if (cur >= 0) {
#pragma unroll
                for (int v = 0; v < V; ++v) st_f4(reinterpret_cast<float*>(sum_b + (cur & ~1)) + 4 * L * v, acc[v]);
              }
              cur = dsto[u];
#pragma unroll
              for (int v = 0; v < V; ++v)                  // every slot has one writer per pass: start from zero
                acc[v] = (cur & 1) ? ld_f4(reinterpret_cast<const float*>(sum_b + (cur & ~1)) + 4 * L * v)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = fma4_packed(wv[u], xv[u][v], acc[v]);
          }
        }
        if (cur >= 0) {
#pragma unroll
          for (int v = 0; v < V; ++v) st_f4(reinterpret_cast<float*>(sum_b + (cur & ~1)) + 4 * L * v, acc[v]);
        }
        __syncwarp();
#pragma unroll
        for (int q = 1; q < NSUBW; ++q) {                  // carries, in sub-warp order
          const int cr = s_carry_row[q];
          if (cr >= 0)
            for (int i = 4 * lane; i < LD; i += 128)
              st_f4(s_sum + cr * LD + i, add4(ld_f4(s_sum + cr * LD + i), ld_f4(s_carry + q * LD + i)));
        }
      }
    }
    __syncwarp();
    LGC_PHASE(1)

    // ---- epilogue in shared memory (same arithmetic as epi_finish), in place
    if (n_in) mbar_wait_parity(bar0 + 8u * st, par);
    LGC_PHASE(2)
#pragma unroll
    for (int j = 0; j < RPS; ++j) {
      const int t = j * NSUBW + sw;
      if (!((okbits >> t) & 1u)) continue;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int off = t * LD + 4 * (sl + L * v);
        const float4 s = ((zbits >> t) & 1u) ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_f4(s_sum + off);
        if (MODE == EPI_PLAIN) {
          float4 r = make_float4(args.scale * s.x, args.scale * s.y, args.scale * s.z, args.scale * s.w);
          if (args.addend) {
            const float4 q = ld_f4(buf0 + off);
            r.x = fmaf(args.beta, q.x, r.x); r.y = fmaf(args.beta, q.y, r.y);
            r.z = fmaf(args.beta, q.z, r.z); r.w = fmaf(args.beta, q.w, r.w);
          }
          st_f4(buf0 + off, r);
        } else if (MODE == EPI_FWD_INIT) {
          const float4 q = ld_f4(buf0 + off);
          if (args.y) st_f4(buf1 + off, s);
          float4 r;                              // out = x * alpha0; out = out + x1 * alpha1
          r.x = __fadd_rn(__fmul_rn(q.x, args.a0), __fmul_rn(s.x, args.a1));
          r.y = __fadd_rn(__fmul_rn(q.y, args.a0), __fmul_rn(s.y, args.a1));
          r.z = __fadd_rn(__fmul_rn(q.z, args.a0), __fmul_rn(s.z, args.a1));
          r.w = __fadd_rn(__fmul_rn(q.w, args.a0), __fmul_rn(s.w, args.a1));
          st_f4(buf0 + off, r);
        } else if (MODE == EPI_FWD_RMW) {
          float4 o = ld_f4(buf0 + off);
          if (args.y) st_f4(buf1 + off, s);
          o.x = __fadd_rn(o.x, __fmul_rn(s.x, args.a1)); o.y = __fadd_rn(o.y, __fmul_rn(s.y, args.a1));
          o.z = __fadd_rn(o.z, __fmul_rn(s.z, args.a1)); o.w = __fadd_rn(o.w, __fmul_rn(s.w, args.a1));
          st_f4(buf0 + off, o);
        } else if (MODE == EPI_FWD_FINAL) {
          float4 q[kMaxHist];
#pragma unroll
          for (int i = 0; i < kMaxHist; ++i)
            if (i < args.n_hist) q[i] = ld_f4(buf0 + i * TF + off);
          st_f4(buf0 + off, final_mean(q, args, s));
        } else {  // EPI_ADAM
          const float4 z = ld_f4(buf0 + off);
          float4 p = ld_f4(buf1 + off), m = ld_f4(buf2 + off), vv = ld_f4(buf3 + off);
          const AdamScalars ad = args.adam_dev ? *args.adam_dev : args.adam;
          const float ib = __frcp_rn(ad.bc2_sqrt);
          adam_update_fast(p.x, m.x, vv.x, fmaf(args.scale, s.x, z.x), ad, ib);
          adam_update_fast(p.y, m.y, vv.y, fmaf(args.scale, s.y, z.y), ad, ib);
          adam_update_fast(p.z, m.z, vv.z, fmaf(args.scale, s.z, z.z), ad, ib);
          adam_update_fast(p.w, m.w, vv.w, fmaf(args.scale, s.w, z.w), ad, ib);
          st_f4(buf1 + off, p);
          st_f4(buf2 + off, m);
          st_f4(buf3 + off, vv);
        }
      }
    }

    // ---- result tiles: shared memory -> HBM, asynchronously; then refill this ring stage
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    LGC_PHASE(3)
    if (lane == 0) {
      // Only rows this kernel owns are written: the hub rows of the tile belong to the heavy-row
      // kernels, which may run CONCURRENTLY on a second stream. A tile without hub rows (the common
      // case) leaves as one copy per output table, otherwise one copy per run of light rows.
      auto store_rows = [&](int a, int n) {
        const uint32_t bytes = (uint32_t)n * LD * 4;
        const size_t goff = (size_t)(row0 + a) * LD;
        const int so = a * LD;
        if (MODE == EPI_PLAIN) {
          bulk_store(args.y + goff, buf0 + so, bytes, pol);
        } else if (MODE == EPI_ADAM) {
          bulk_store(args.p + goff, buf1 + so, bytes, pol);
          bulk_store(args.m + goff, buf2 + so, bytes, pol);
          bulk_store(args.v + goff, buf3 + so, bytes, pol);
        } else {
          bulk_store(args.acc + goff, buf0 + so, bytes, pol);
          if (args.y) bulk_store(args.y + goff, buf1 + so, bytes, pol);
        }
      };
      const unsigned full = rows_here >= 32 ? 0xffffffffu : ((1u << rows_here) - 1u);
      if (okbits == full) {
        store_rows(0, rows_here);
      } else {
        unsigned todo = okbits;
        while (todo) {
          const int a = __ffs(todo) - 1;
          const unsigned rest = ~(todo >> a);
          const int n = rest ? __ffs(rest) - 1 : 32 - a;
          todo = (n >= 32) ? 0u : (todo & ~(((1u << n) - 1u) << a));
          store_rows(a, n);
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      request_csr(tile + nw, st ^ 1, k + 1);                              // next tile's CSR slice
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // stage readable again
      request_tile(tile + 2 * nw, st);
    }
    __syncwarp();
    LGC_PHASE(4)
#ifdef LGC_PHASE_CLOCKS
    if (phases && lane == 0) s_phase[5] += 1;
#endif
  }
#ifdef LGC_PHASE_CLOCKS
  if (phases && lane == 0)
    for (int i = 0; i < 6; ++i) atomicAdd(phases + i, s_phase[i]);
#endif
#undef LGC_PHASE
}

// ---------------------------------------------------------------------------------- heavy rows
// Hub rows (mostly items): every chunk (<= 256 source-sorted edges of one row) is a work item; a
// warp walks its chunks (c, c + #warps, ...) as ONE continuous stream of groups of R edges, so the
// copy pipeline never drains between chunks. Neighbour rows are copied with cp.async (LDGSTS: no
// destination register, no in-order wait) into a per-warp shared-memory ring of kHeavyStages
// groups, kHeavyStages-1 groups ahead of the FMAs (~12 KB per warp in flight). The kernel is bound
// by issued instructions (profiles/r1f_light_bottleneck.md), hence: one coalesced load of a
// group's (src, w) by R lanes and shuffles instead of per-edge broadcast loads, weights transposed
// in shared memory so that a stream reads its EPL weights as vectors, packed FFMA2, and no padding
// guards (the ring is zeroed once; rows that are not copied keep finite data and get weight 0).
constexpr int kHeavyWarps = 4;         // warps per CTA (independent: no CTA barrier)
constexpr int kHeavyStages = 4;
constexpr int kHeavyQueue = 8;         // chunk descriptors between the fetch cursor and the FMAs
#ifndef LGC_HEAVY_GROUP_BYTES
#define LGC_HEAVY_GROUP_BYTES 4096
#endif
constexpr int kHeavyGroupBytes = LGC_HEAVY_GROUP_BYTES;   // rows of one ring slot

template <int L, int V>
struct HeavyCfg {
  static constexpr int LD = 4 * L * V;
  static constexpr int RPW = 32 / L;                                   // edge streams per warp
  static constexpr int R0 = kHeavyGroupBytes / (LD * 4) > 32 ? 32 : kHeavyGroupBytes / (LD * 4);   // <= one lane per edge
  static constexpr int RMIN = 2 * RPW > 32 ? 32 : 2 * RPW;
  static constexpr int R = R0 / RPW * RPW < RMIN ? RMIN : R0 / RPW * RPW;   // rows per group (multiple of RPW)
  static constexpr int EPL = R / RPW;                                  // edges per lane-stream and group
  static constexpr size_t GROUP_BYTES = ((size_t)R * LD * 4 + (size_t)R * 4 + 15) / 16 * 16;   // rows + their weights
  static constexpr size_t WARP_BYTES = kHeavyStages * GROUP_BYTES + kHeavyQueue * 16;
  static constexpr size_t SMEM = kHeavyWarps * WARP_BYTES + 16;
  static_assert(R <= 32, "one lane per edge of a group");
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc) : "memory");
}

template <int L, int V, int MODE>
__global__ void __launch_bounds__(32 * kHeavyWarps)
k_spmm_heavy(const int4* __restrict__ chunks, int num_chunks, const int32_t* __restrict__ src,
             const float* __restrict__ w, const float* __restrict__ x, float* __restrict__ partials,
             EpiArgs args) {
  using C = HeavyCfg<L, V>;
  constexpr int LD = C::LD, RPW = C::RPW, R = C::R, EPL = C::EPL, S = kHeavyStages;
  extern __shared__ __align__(16) uint8_t smem_heavy[];
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbase = smem_heavy + (size_t)wic * C::WARP_BYTES;
  auto slot_rows = [&](int sl_) { return reinterpret_cast<float*>(wbase + (size_t)sl_ * C::GROUP_BYTES); };
  auto slot_w = [&](int sl_) { return slot_rows(sl_) + R * LD; };       // transposed: [stream][EPL]
  int4* const queue = reinterpret_cast<int4*>(wbase + (size_t)S * C::GROUP_BYTES);
  const int sub = lane / L, sl = lane % L;
  const int n_warps = gridDim.x * kHeavyWarps;
  const int first = blockIdx.x * kHeavyWarps + wic;
  if (first >= num_chunks) return;

  for (int i = 16 * lane; i < (int)(S * C::GROUP_BYTES); i += 512)       // ring: finite data everywhere
    *reinterpret_cast<float4*>(wbase + i) = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- fetch cursor: group (f_k, f_g) of the warp's chunk sequence; `nxt` = descriptor of chunk f_k + 1
  int f_k = 0, f_g = 0, f_beg, f_end, f_ng;
  bool f_valid = true;
  int4 nxt;
  {
    const int4 c = chunks[first];
    f_beg = c.y; f_end = c.z; f_ng = (c.z - c.y + R - 1) / R;
    if (lane == 0) queue[0] = c;
    const int c1 = first + n_warps;
    nxt = c1 < num_chunks ? chunks[c1] : make_int4(0, 0, 0, 0);
  }
  int my_idx, my_w_bits;                            // this lane's edge of the group fetched last
  auto fetch = [&]() {
    my_idx = -1; my_w_bits = 0;
    if (!f_valid) return;
    const int e = f_beg + f_g * R + lane;
    if (lane < R && e < f_end) { my_idx = src[e]; my_w_bits = __float_as_int(w[e]); }
    if (++f_g == f_ng) {                            // next chunk: its descriptor is already in registers
      ++f_k; f_g = 0;
      const int c_idx = first + f_k * n_warps;
      f_valid = c_idx < num_chunks;
      if (f_valid) {
        f_beg = nxt.y; f_end = nxt.z; f_ng = (nxt.z - nxt.y + R - 1) / R;
        if (lane == 0) queue[f_k % kHeavyQueue] = nxt;
        const int c2 = c_idx + n_warps;
        if (c2 < num_chunks) nxt = chunks[c2];
      }
    }
  };
  auto issue = [&](int slot_) {                     // the group held in (my_idx, my_w_bits) -> ring slot
    float* rows = slot_rows(slot_);
    if (lane < R) slot_w(slot_)[(lane % RPW) * EPL + lane / RPW] = __int_as_float(my_w_bits);
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int r = i * RPW + sub;
      const int idx = __shfl_sync(0xffffffffu, my_idx, r);
      if (idx >= 0) {
        const float* xr = x + (size_t)idx * LD + 4 * sl;
#pragma unroll
        for (int v = 0; v < V; ++v) cp_async16(rows + r * LD + 4 * (sl + L * v), xr + 4 * L * v);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per step, empty or not
  };

  __syncwarp();
#pragma unroll 1
  for (int g = 0; g < S - 1; ++g) { fetch(); issue(g); }
  fetch();

  // ---- consumer: group (c_k, c_g)
  int c_k = 0, c_g = 0;
  int4 cur = chunks[first];
  int c_ng = (cur.z - cur.y + R - 1) / R;
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int t = 0;; ++t) {
    issue((t + S - 1) % S);                         // into the slot consumed in the previous iteration
    fetch();                                        // indices one group ahead of their issue
    asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
    __syncwarp();                                   // every lane's copies of group t have landed
    const float* rows = slot_rows(t % S);
    const float* ws = slot_w(t % S) + sub * EPL;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int r = i * RPW + sub;
      const float wv = ws[i];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = fma4_packed(wv, ld_f4(rows + r * LD + 4 * (sl + L * v)), acc[v]);
    }
    __syncwarp();                                   // slot t % S may be refilled
    if (++c_g < c_ng) continue;
    // ---- end of a chunk: combine the RPW edge streams (fixed order: deterministic), write, next chunk
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
    if (sub == 0) {
      if (cur.w >= 0) {
        float* pr = partials + (size_t)cur.w * LD + 4 * sl;
#pragma unroll
        for (int v = 0; v < V; ++v) st_f4(pr + 4 * L * v, acc[v]);
      } else {
        const size_t off = (size_t)cur.x * LD + 4 * sl;
#pragma unroll
        for (int v = 0; v < V; ++v) epilogue<MODE>(args, off + 4 * L * v, acc[v]);
      }
    }
    ++c_k; c_g = 0;
    if (first + c_k * n_warps >= num_chunks) break;
    cur = queue[c_k % kHeavyQueue];
    c_ng = (cur.z - cur.y + R - 1) / R;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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

// The light-row kernel is bound by issued instructions and leaves HBM half idle; the heavy-row
// kernel waits on random 256-byte reads and leaves the issue slots idle. They touch disjoint rows,
// so (except for the ADAM epilogue, whose tiles fill the shared memory) they CAN run concurrently
// (LGC_SPMM_OVERLAP=1): heavy + finish on a side stream between a fork and a join event, with the persistent grids sized
// so that `kHeavyCtasOverlap` heavy CTAs and the light CTAs fit one SM together. Works inside
// stream capture (the side stream joins the capture through the events).
#ifndef LGC_HEAVY_CTAS_OVERLAP
#define LGC_HEAVY_CTAS_OVERLAP 1
#endif
constexpr int kHeavyCtasOverlap = LGC_HEAVY_CTAS_OVERLAP;
constexpr size_t kSmemPerSM = 228 * 1024, kSmemPerCtaReserved = 1024;
struct OverlapCtx {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
OverlapCtx* overlap_ctx() {
  static OverlapCtx ctx[64];
  // Off by default: measured at c2 (profiles/r1f_light_bottleneck.md) both kernels are bound by issued
  // instructions, so running them side by side gains nothing (3.61 vs 3.54 ms per step).
  static const bool enabled = [] { const char* e = getenv("LGC_SPMM_OVERLAP"); return e && atoi(e) != 0; }();
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  OverlapCtx& c = ctx[dev];
  if (!c.ok) {
    // created on the first (eager) launch: callers that capture CUDA graphs run eager steps first
    if (cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    c.ok = true;
  }
  return &c;
}

// L, V: row geometry of the heavy-row kernels (widest sub-warp); LL, LV: of the light-row kernel
// (few lanes per row: more rows per warp instruction, more gathers in flight per lane)
template <int L, int V, int LL, int LV, int MODE>
int launch_lv(const lgc_graph* g, const float* x, const EpiArgs& a, float* partials, cudaStream_t st) {
  const int threads = 256;
  const int64_t n = g->num_nodes;
  using LC = LightCfg<LL, LV, MODE>;
  using HC = HeavyCfg<L, V>;
  const int nbuf = MODE == EPI_FWD_FINAL ? a.n_hist : LC::NBUF;
  int warps = LC::WARPS;
  while (warps > 1 && LC::smem(nbuf, warps) > kMaxDynamicSmem) --warps;
  const size_t smem = LC::smem(nbuf, warps);
  if (smem > kMaxDynamicSmem) {
    set_error("row width x operand tiles do not fit in shared memory (ld=" + std::to_string(LC::LD) + ")");
    return LGC_ERR_UNSUPPORTED;
  }
  // light CTAs per SM that leave room for the overlapping heavy CTAs
  const size_t heavy_need = kHeavyCtasOverlap * (HC::SMEM + kSmemPerCtaReserved);
  const int light_cap = heavy_need < kSmemPerSM ? (int)((kSmemPerSM - heavy_need) / (smem + kSmemPerCtaReserved)) : 0;
  OverlapCtx* ov = (MODE != EPI_ADAM && g->num_chunks > 0 && light_cap >= 1 && !rows_kernel_enabled() && !sweep_get(g, HC::LD)) ? overlap_ctx() : nullptr;
  cudaStream_t hs = st;
  if (ov) {
    LGC_CUDA(cudaEventRecord(ov->fork, st));
    LGC_CUDA(cudaStreamWaitEvent(ov->side, ov->fork, 0));
    hs = ov->side;
  }
  // ---- high-degree rows: the sweep kernel (sweep.cu) when a schedule exists for this row width ...
  const SweepSched* sw = (g->num_chunks > 0 || rows_kernel_enabled()) ? sweep_get(g, HC::LD) : nullptr;
  if (sw && sweep_has_rows(sw)) {
    int rc = launch_sweep(g, sw, HC::LD, x, (EpiMode)MODE, a, partials, st);
    if (rc) return rc;
    int64_t n_split = 0;
    const int4* split = sweep_split_rows(sw, &n_split);
    if (n_split > 0) {
      {
        ProfScope ps(PROF_FINISH + (MODE & 3), st);
        k_spmm_finish<L, V, MODE><<<(int)n_split, threads, 0, st>>>(split, (int)n_split, partials, a);
      }
      LGC_LAUNCH_CHECK();
    }
  }
  // ---- ... else the chunked heavy-row kernel (+ the sums of split rows): launched first so that
  // their CTAs find room
  if (!sw && g->num_chunks > 0) {
    static int occ_heavy_dev[kMaxDevices] = {};   // per instantiation and device: resident CTAs per SM
    int& occ_heavy = occ_heavy_dev[current_device_slot()];
    if (!occ_heavy) {
      LGC_CUDA(cudaFuncSetAttribute(k_spmm_heavy<L, V, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)HC::SMEM));
      int occ = 0;
      LGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmm_heavy<L, V, MODE>, 32 * kHeavyWarps,
                                                             HC::SMEM));
      occ_heavy = occ > 0 ? occ : 1;
    }
    const int per_sm = ov ? std::min(occ_heavy, kHeavyCtasOverlap) : occ_heavy;
    const int grid_heavy = (int)std::min<int64_t>(ceil_div(g->num_chunks, kHeavyWarps), (int64_t)device_sm_count() * per_sm);
    {
      ProfScope ps(PROF_HEAVY + (MODE & 3), hs);
      k_spmm_heavy<L, V, MODE><<<grid_heavy, 32 * kHeavyWarps, HC::SMEM, hs>>>(
          g->chunks, (int)g->num_chunks, g->hsrc, g->hw, x, partials, a);
    }
    LGC_LAUNCH_CHECK();
    if (g->num_split_rows > 0) {
      const int grid_fin = (int)g->num_split_rows;
      {
        ProfScope ps(PROF_FINISH + (MODE & 3), hs);
        k_spmm_finish<L, V, MODE><<<grid_fin, threads, 0, hs>>>(g->split_rows, (int)g->num_split_rows,
                                                                 partials, a);
      }
      LGC_LAUNCH_CHECK();
    }
  }
  if (ov) LGC_CUDA(cudaEventRecord(ov->join, hs));
  // ---- all other rows: the rows kernel (rows.cu) over the schedule's row plan ...
  if (sw && sweep_row_plan(sw)) return launch_rows(g, sweep_row_plan(sw), HC::LD, x, (EpiMode)MODE, a, st);
  // ---- ... or the round-1 light-row kernel (rows up to light_max_degree)
  {
    // per instantiation, device and buffer count: resident CTAs per SM; largest shared-memory size set
    static int occ_light_dev[kMaxDevices][kMaxHist + 1] = {};
    static size_t smem_set_dev[kMaxDevices] = {};
    int* const occ_light = occ_light_dev[current_device_slot()];
    size_t& smem_set = smem_set_dev[current_device_slot()];
    if (smem > smem_set) {
      LGC_CUDA(cudaFuncSetAttribute(k_spmm_light<LL, LV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      smem_set = smem;
    }
    if (!occ_light[nbuf]) {
      int occ = 0;
      LGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmm_light<LL, LV, MODE>, 32 * warps, smem));
      occ_light[nbuf] = occ > 0 ? occ : 1;
    }
    const int per_sm = ov ? std::min(occ_light[nbuf], light_cap) : occ_light[nbuf];
    const int n_tiles = (int)ceil_div(n, LC::TR);
    const int grid = (int)std::min<int64_t>((int64_t)device_sm_count() * per_sm, ceil_div(n_tiles, warps));
    ProfScope ps(PROF_LIGHT + (MODE & 3), st);
    k_spmm_light<LL, LV, MODE><<<grid, 32 * warps, smem, st>>>(g->rowptr, g->src, g->w_hat, x, (int)n,
                                                                g->light_max_degree, light_phase_buffer(), a);
  }
  LGC_LAUNCH_CHECK();
  if (ov) LGC_CUDA(cudaStreamWaitEvent(st, ov->join, 0));
  return LGC_OK;
}

template <int L, int V, int LL, int LV>
int launch_mode(const lgc_graph* g, const float* x, EpiMode mode, const EpiArgs& a, float* partials,
                cudaStream_t st) {
  switch (mode) {
    case EPI_PLAIN: return launch_lv<L, V, LL, LV, EPI_PLAIN>(g, x, a, partials, st);
    case EPI_FWD_INIT: return launch_lv<L, V, LL, LV, EPI_FWD_INIT>(g, x, a, partials, st);
    case EPI_FWD_RMW: return launch_lv<L, V, LL, LV, EPI_FWD_RMW>(g, x, a, partials, st);
    case EPI_ADAM: return launch_lv<L, V, LL, LV, EPI_ADAM>(g, x, a, partials, st);
    case EPI_FWD_FINAL: return launch_lv<L, V, LL, LV, EPI_FWD_FINAL>(g, x, a, partials, st);
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
  const SweepSched* sw = (g->num_chunks > 0 || rows_kernel_enabled()) ? sweep_get(g, ld) : nullptr;   // builds the schedule on first use
  const size_t slots = sw ? sweep_partial_slots(sw) : (size_t)g->num_partial_slots;
  return slots * (size_t)ld;
}

int launch_spmm(const lgc_graph* g, int ld, const float* x, EpiMode mode, const EpiArgs& a,
                float* partials, cudaStream_t st) {
  RowShape rs;
  if (!row_shape(ld, &rs)) {
    set_error("unsupported row width ld=" + std::to_string(ld));
    return LGC_ERR_UNSUPPORTED;
  }
#define LGC_CASE(HL, HV, LL, LV) \
  if (rs.L == HL && rs.V == HV) return launch_mode<HL, HV, LL, LV>(g, x, mode, a, partials, st);
  // light-row geometry == heavy-row geometry: narrower sub-warps (4 lanes x 4 float4) measured
  // 1.5x slower -- 8 rows per warp instruction hit the same shared-memory banks in the epilogue
  LGC_CASE(16, 1, 8, 2) LGC_CASE(16, 2, 16, 2) LGC_CASE(16, 3, 16, 3) LGC_CASE(16, 4, 16, 4)
  LGC_CASE(8, 1, 8, 1) LGC_CASE(8, 3, 8, 3) LGC_CASE(8, 5, 8, 5)
  LGC_CASE(4, 1, 4, 1) LGC_CASE(4, 3, 4, 3) LGC_CASE(4, 5, 4, 5)
  LGC_CASE(2, 1, 2, 1) LGC_CASE(1, 1, 1, 1)
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

extern "C" int lgc_graph_plan_info(const lgc_graph_t* g, int ld, lgc_plan_info* info) {
  LGC_REQUIRE(g && info, "null argument");
  *info = lgc_plan_info{};
  const SweepSched* sw = (g->num_chunks > 0 || rows_kernel_enabled()) ? sweep_get(g, ld) : nullptr;
  if (sw) sweep_info(sw, info);
  info->has_plan = sw ? 1 : 0;
  return LGC_OK;
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

static int epilogue_args(const lgc_spmm_epilogue* e, const float* x, EpiArgs& a) {
  a.y = e->y; a.acc = e->acc; a.xrow = e->xrow; a.addend = e->addend;
  a.a0 = e->a0; a.a1 = e->a1; a.scale = e->scale; a.beta = e->beta;
  a.p = e->p; a.m = e->m; a.v = e->v;
  switch (e->mode) {
    case LGC_EPI_PLAIN: LGC_REQUIRE(e->y && e->y != x, "PLAIN needs y (not aliasing x)"); break;
    case LGC_EPI_FWD_INIT: LGC_REQUIRE(e->acc && e->xrow, "FWD_INIT needs acc and xrow"); break;
    case LGC_EPI_FWD_RMW: LGC_REQUIRE(e->acc, "FWD_RMW needs acc"); break;
    case LGC_EPI_ADAM:
      LGC_REQUIRE(e->addend && e->p && e->m && e->v && (e->step >= 1 || e->adam_scalars),
                  "ADAM needs addend, p, m, v and step >= 1 (or device scalars)");
      if (e->adam_scalars) a.adam_dev = reinterpret_cast<const AdamScalars*>(e->adam_scalars);
      else a.adam = make_adam_scalars(e->lr, e->beta1, e->beta2, e->eps, e->step);
      break;
    case LGC_EPI_FWD_FINAL:
      LGC_REQUIRE(e->acc && e->n_hist >= 1 && e->n_hist <= kMaxHist, "FWD_FINAL needs acc and 1..6 layer tables");
      a.n_hist = e->n_hist;
      for (int i = 0; i < e->n_hist; ++i) {
        LGC_REQUIRE(e->hist[i], "FWD_FINAL: null layer table");
        a.hist[i] = e->hist[i];
        a.ah[i] = e->ah[i];
      }
      break;
    default: LGC_REQUIRE(false, "unknown epilogue mode");
  }
  return LGC_OK;
}

extern "C" int lgc_spmm_ex(const lgc_graph_t* g, int ld, const float* x, const lgc_spmm_epilogue* e,
                           void* workspace, size_t workspace_bytes, void* stream) {
  LGC_REQUIRE(g && x && e, "null argument");
  if (workspace_bytes < lgc_spmm_workspace_bytes(g, ld)) {
    set_error("lgc_spmm_ex: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  EpiArgs a;
  int rc = epilogue_args(e, x, a);
  if (rc) return rc;
  return launch_spmm(g, ld, x, (EpiMode)e->mode, a, (float*)workspace, (cudaStream_t)stream);
}

namespace lgc {
namespace {
// the fused epilogue on existing row sums: one thread per float4 column of a row
template <int MODE>
__global__ void k_epilogue_apply(int64_t n_vec, const float* __restrict__ sums, EpiArgs args) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x)
    epilogue<MODE>(args, (size_t)i * 4, ld_f4(sums + i * 4));
}
}  // namespace
}  // namespace lgc

extern "C" int lgc_epilogue_apply(int64_t n_rows, int ld, const float* sums, const lgc_spmm_epilogue* e,
                                  void* stream) {
  LGC_REQUIRE(sums && e && n_rows >= 0 && ld > 0 && ld % 4 == 0, "bad argument");
  EpiArgs a;
  int rc = epilogue_args(e, nullptr, a);
  if (rc) return rc;
  const int64_t n_vec = n_rows * (ld / 4);
  if (n_vec == 0) return LGC_OK;
  const int grid = (int)std::min<int64_t>(ceil_div(n_vec, 256), (int64_t)device_sm_count() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (e->mode) {
    case LGC_EPI_PLAIN: k_epilogue_apply<EPI_PLAIN><<<grid, 256, 0, st>>>(n_vec, sums, a); break;
    case LGC_EPI_FWD_INIT: k_epilogue_apply<EPI_FWD_INIT><<<grid, 256, 0, st>>>(n_vec, sums, a); break;
    case LGC_EPI_FWD_RMW: k_epilogue_apply<EPI_FWD_RMW><<<grid, 256, 0, st>>>(n_vec, sums, a); break;
    case LGC_EPI_ADAM: k_epilogue_apply<EPI_ADAM><<<grid, 256, 0, st>>>(n_vec, sums, a); break;
    case LGC_EPI_FWD_FINAL: k_epilogue_apply<EPI_FWD_FINAL><<<grid, 256, 0, st>>>(n_vec, sums, a); break;
    default: break;
  }
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

// Forward chain shared by lgc_propagate and lgc_train_step: layers 1..K-1 store x_l = A x_{l-1}
// (plain epilogue), layer K folds the whole mean sum_l alpha_l x_l into its epilogue (FWD_FINAL).
// `xs` holds K-1 tables. Row streams: (K-1) + (K+1) instead of the running sum's 3K-1.
namespace lgc {
int propagate_chain(const lgc_graph* g, int ld, int K, const float* alpha, const float* x0, float* out,
                    float* const* xs, float* partials, cudaStream_t st) {
  const float* cur = x0;
  for (int l = 1; l < K; ++l) {
    EpiArgs a;
    a.y = xs[l - 1];
    int rc = launch_spmm(g, ld, cur, EPI_PLAIN, a, partials, st);
    if (rc) return rc;
    cur = xs[l - 1];
  }
  EpiArgs a;
  a.acc = out;
  a.a1 = alpha[K];
  a.n_hist = K;
  a.hist[0] = x0; a.ah[0] = alpha[0];
  for (int l = 1; l < K; ++l) { a.hist[l] = xs[l - 1]; a.ah[l] = alpha[l]; }
  return launch_spmm(g, ld, cur, EPI_FWD_FINAL, a, partials, st);
}
}  // namespace lgc

extern "C" size_t lgc_propagate_workspace_bytes(const lgc_graph_t* g, int ld, int num_layers) {
  if (!g) return 0;
  size_t t = (size_t)g->num_nodes * ld;
  size_t n_tmp = num_layers > 1 ? (size_t)num_layers - 1 : 0;
  return (n_tmp * t + spmm_partials_floats(g, ld)) * sizeof(float);
}

extern "C" int lgc_propagate(const lgc_graph_t* g, int ld, int num_layers, const float* h_alpha,
                             const float* x0, float* out, void* workspace, size_t workspace_bytes,
                             void* stream) {
  LGC_REQUIRE(g && h_alpha && x0 && out, "null argument");
  LGC_REQUIRE(num_layers >= 0 && num_layers <= kMaxHist, "num_layers must be in 0..6");
  LGC_REQUIRE(x0 != out, "x0 and out must not alias");
  if (workspace_bytes < lgc_propagate_workspace_bytes(g, ld, num_layers)) {
    set_error("lgc_propagate: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  NvtxRange nvtx("lgc_propagate");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t t = (size_t)g->num_nodes * ld;
  if (num_layers == 0) {
    k_scale<<<kNumSMs * 8, 256, 0, st>>>((const float4*)x0, (float4*)out, h_alpha[0], (int64_t)(t / 4));
    LGC_LAUNCH_CHECK();
    return LGC_OK;
  }
  float* ws = (float*)workspace;
  float* xs[kMaxHist] = {};
  for (int l = 0; l + 1 < num_layers; ++l) xs[l] = ws + (size_t)l * t;
  float* partials = ws + (size_t)(num_layers - 1) * t;
  return propagate_chain(g, ld, num_layers, h_alpha, x0, out, xs, partials, st);
}

// Diagnostics: cycles per phase of k_spmm_light since the last call (needs LGC_LIGHT_PHASES=1 in the
// environment; otherwise all zeros). out[0..4] = wait CSR | gathers | wait operands | epilogue |
// lane-0 store/refill, out[5] = warp-tiles.
extern "C" int lgc_debug_light_phases(unsigned long long* out8) {
  unsigned long long* buf = light_phase_buffer();
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  if (!buf) return LGC_OK;
  LGC_CUDA(cudaDeviceSynchronize());
  LGC_CUDA(cudaMemcpy(out8, buf, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  LGC_CUDA(cudaMemset(buf, 0, 8 * sizeof(unsigned long long)));
  return LGC_OK;
}
