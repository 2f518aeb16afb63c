// recommendK scoring: top-k items per user from the final embeddings
// (replaces `src[user_id_list] @ dst.t()` + `.cpu()` + `pred * (1 - mask)` + `.topk(k)` of
// reference src/lightgcn.py:169-182; K10/K11 of SURVEY.md 2.3).
//
// The [U, I] score matrix never reaches HBM. Pipeline per chunk of users (all on one stream):
//   prep     fp32 tables -> fp16 copies, scaled by a power of two per table (exact scaling), plus
//            row norms (error bound)                                             k_absmax/k_convert
//   gemm     tcgen05.mma (kind::f16, fp32 accumulators in TMEM), operands staged by TMA with the
//            128-byte swizzle; persistent, one CTA per SM: 256 users (2 x M=128) x 128 items per
//            step; 8 epilogue warps read the accumulators with tcgen05.ld and keep only the
//            maximum of every 16 consecutive items (+ the maximum of the 128-item tile) k_score_gemm
//   thresh   T_u = (k + n_seen_u)-th largest tile maximum: a LOWER bound (up to the proven fp16
//            error eps_u) of the k-th best masked score, because at most n_seen_u of those tiles
//            owe their maximum to a seen item                                          k_threshold
//   rescore  every 16-item group whose maximum reaches T_u - 2 eps_u is re-scored exactly in fp32
//            with the multiplicative seen-mask; items >= T_u - eps_u become candidates   k_rescore
//   select   exact top-k of candidates + seen items (masked score 0.0), ordered by
//            (score desc, item asc)                                                      k_select
//   fallback users whose bound cannot be proven (k + n_seen > #tiles, candidate overflow) are
//            scored exhaustively in fp32                                             k_exhaustive
// Results equal the fp32 reference apart from exact ties / fp32 summation-order near-ties.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace lgc {
namespace {

constexpr int kUserBlock = 256;   // users per CTA step (two M=128 halves)
constexpr int kTileN = 128;       // items per MMA tile (N)
constexpr int kGroup = 16;        // items per stored maximum
constexpr int kCap = 128;         // exact candidates kept per user before falling back
constexpr int kGemmThreads = 320; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr float kEpsRel = 0.0025f;  // > fp16 input rounding (2 * 2^-11) + fp16 output rounding (2^-11)
                                    //   + tensor-core fp32 accumulation slack, relative to |a||b|
constexpr size_t kWorkspaceBudget = 16ull << 30;

struct ScoreScalars {
  unsigned amax_bits, bmax_bits, bnorm_bits;
  int fb_count;
  unsigned long long n_groups, n_emitted;
};

__host__ __device__ inline int round_up(int64_t x, int64_t m) { return (int)((x + m - 1) / m * m); }

__device__ __forceinline__ int scale_exp(unsigned max_bits) {
  const float m = __uint_as_float(max_bits);
  if (!(m > 0.f) || isinf(m) || isnan(m)) return 0;
  return 14 - ilogbf(m);            // max |x| * 2^e lands in [2^14, 2^15): finite in fp16
}
__host__ __device__ inline int out_scale_exp(int kp) { return kp <= 64 ? -22 : (kp <= 128 ? -23 : -24); }

// ------------------------------------------------------------------------------------ prep
__global__ void k_absmax(const float* __restrict__ tab, int ld, int d, const int64_t* __restrict__ ids,
                         int64_t n_rows, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const float* row = tab + (size_t)(ids ? ids[r] : r) * ld;
    for (int c = lane; c < d; c += 32) m = fmaxf(m, fabsf(row[c]));
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

// One warp per row: fp16(x * 2^e), zero padded to kp columns; rows >= n_rows are zero.
__global__ void k_convert(const float* __restrict__ tab, int ld, int d, const int64_t* __restrict__ ids,
                          int64_t id_offset, int64_t n_rows, int64_t n_rows_pad, int kp,
                          const unsigned* __restrict__ max_bits, __half* __restrict__ out,
                          float* __restrict__ norm_out, unsigned* __restrict__ norm_max_bits) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows_pad) return;
  const int e = scale_exp(*max_bits);
  const bool live = r < n_rows;
  const float* row = live ? tab + (size_t)(ids ? ids[id_offset + r] : id_offset + r) * ld : nullptr;
  float ss = 0.f;
  for (int c = lane; c < kp; c += 32) {
    float v = (live && c < d) ? scalbnf(row[c], e) : 0.f;
    ss = fmaf(v, v, ss);
    out[(size_t)r * kp + c] = __float2half_rn(v);
  }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) {
    const float nrm = sqrtf(ss) * 1.0001f;        // rounded up: it feeds an error BOUND
    if (norm_out) norm_out[r] = nrm;
    if (norm_max_bits && nrm > 0.f) atomicMax(norm_max_bits, __float_as_uint(nrm));
  }
}

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M=128, K=16 fp16, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
#define LGC_R8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define LGC_RW8(v, o) "+r"(v[o]), "+r"(v[o + 1]), "+r"(v[o + 2]), "+r"(v[o + 3]), "+r"(v[o + 4]), "+r"(v[o + 5]), "+r"(v[o + 6]), "+r"(v[o + 7])
// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : LGC_R8(v, 0), LGC_R8(v, 8), LGC_R8(v, 16), LGC_R8(v, 24)
      : "r"(taddr)
      : "memory");
}
// wait for the outstanding tcgen05.ld of two 32-column chunks (registers threaded through)
__device__ __forceinline__ void tc_wait_ld64(uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : LGC_RW8(a, 0), LGC_RW8(a, 8), LGC_RW8(a, 16), LGC_RW8(a, 24), LGC_RW8(b, 0), LGC_RW8(b, 8),
                 LGC_RW8(b, 16), LGC_RW8(b, 24)
               :
               : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// K-major operand tile in the canonical 128-byte-swizzle layout: rows of 128 B (64 fp16),
// 8-row swizzle atoms 1024 B apart; descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= (uint64_t)0 << 16;                        // leading byte offset: unused (K extent 32 B < atom)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
                            ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// ------------------------------------------------------------------------------------ GEMM + group max
// 16 words of this thread's TMEM lane, columns taddr.. (registers -> TMEM)
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: A rows live in TMEM lanes, two fp16 per 32-bit column
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// maximum of 16 consecutive scores v[o..o+15]; columns at or beyond n_valid are masked out
__device__ __forceinline__ float group_max16(const uint32_t* v, int n_valid) {
  float f[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) f[x] = __uint_as_float(v[x]);
  if (n_valid < 16) {
#pragma unroll
    for (int x = 0; x < 16; ++x)
      if (x >= n_valid) f[x] = -INFINITY;
  }
  const float m0 = max3(f[0], f[1], f[2]), m1 = max3(f[3], f[4], f[5]), m2 = max3(f[6], f[7], f[8]);
  const float m3 = max3(f[9], f[10], f[11]), m4 = max3(f[12], f[13], f[14]);
  return fmaxf(max3(m0, m1, f[15]), max3(m2, m3, m4));
}

// Persistent, one CTA per SM. Per step: 256 users (two M=128 halves) x 128 items, K = kp.
// TMEM (512 columns): a ring of three 128-column fp32 accumulators (one half-tile each) at columns
// 0/128/256, and -- TS mode -- the users' fp16 rows as the A operand at columns 384.. (kp/2 columns
// per half). With A in TMEM the tensor core streams only B from shared memory: an SS-mode M=128 x
// N=128 step needs A + B = 128 B/clk of shared-memory reads and measured 58 % of peak on its own.
// SS mode (A tiles in shared memory via TMA) remains for kp > 128, where A does not fit in TMEM.
// Epilogue (8 warps, thread = user row): tcgen05.ld 16 columns at a time, maximum of each group of
// 16 items; stored per group as an UPPER bound fp16((max + e) * out_scale) rounded up, and per tile
// as a LOWER bound fp16((max - e) * out_scale) rounded down, e = eps_rel*|a_u|*max|b_tile| + eps_abs.
//   gmax16 : [n_tiles][u_pad][4] uint32 = upper bounds of groups 2p (low half) and 2p+1 of a tile;
//            one 16-byte record per (tile, user): one STG.128 here, one LDG.128 in k_scan
//   gtile  : [n_tiles][u_pad] uint32 = fp16 lower bound (low half) and upper bound (high half) of
//            the tile maximum
template <bool TS, int KATOMS>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_score_gemm(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const __half* __restrict__ a16, const float* __restrict__ anorm, const float* __restrict__ btile,
             int n_blocks, int n_tiles, int n_items, int n_stages, int u_pad, float out_scale,
             float eps_abs, uint32_t* __restrict__ gmax16, uint32_t* __restrict__ gtile, int dbg_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int katoms = KATOMS;
  constexpr uint32_t tile_bytes = 16384u * KATOMS;             // one 128-row operand tile
  uint8_t* smem_a = smem;                                      // SS: 2 tiles; TS: none
  uint8_t* smem_b = smem + (TS ? 0 : 2) * tile_bytes;          // n_stages tiles
  uint64_t* bars = (uint64_t*)(smem_b + (size_t)n_stages * tile_bytes);
  // barrier slots: 0 a_full, 1 a_empty, 2..4 acc_full, 5..7 acc_empty, 8+s b_full, 16+s b_empty
  uint32_t* tmem_slot = (uint32_t*)(bars + 24);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kp = KATOMS * 64;
  const uint32_t a_col0 = 384u;                                // TS: A of half h at a_col0 + h * kp/2

  if (threadIdx.x == 0) {
    mbar_init(bar(0), TS ? 8 : 1); mbar_init(bar(1), 1);
    for (int b = 0; b < 3; ++b) { mbar_init(bar(2 + b), 1); mbar_init(bar(5 + b), 4); }
    for (int s = 0; s < n_stages; ++s) { mbar_init(bar(8 + s), 1); mbar_init(bar(16 + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      int s = 0; uint32_t ph = 0, pa = 0;
      for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        if (!TS) {
          mbar_wait(bar(1), pa ^ 1);
          mbar_expect_tx(bar(0), 2 * tile_bytes);
          for (int h = 0; h < 2; ++h)
            for (int ka = 0; ka < katoms; ++ka)
              tma_load_2d(smem_u32(smem_a + h * tile_bytes + ka * 16384), &map_a, ka * 64,
                          blk * kUserBlock + h * 128, bar(0));
          pa ^= 1;
        }
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait(bar(16 + s), ph ^ 1);
          mbar_expect_tx(bar(8 + s), tile_bytes);
          for (int ka = 0; ka < katoms; ++ka)
            tma_load_2d(smem_u32(smem_b + (size_t)s * tile_bytes + ka * 16384), &map_b, ka * 64, t * kTileN,
                        bar(8 + s));
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) control flow so every address stays
    // in uniform registers; one elected lane issues. Descriptors are base + compile-time offsets.
    int s = 0, buf = 0; uint32_t ph = 0, pa = 0, pb = 0;
    constexpr int ksteps = KATOMS * 4;
    const uint64_t b_desc0 = umma_desc(smem_u32(smem_b));
    const uint64_t a_desc0 = umma_desc(smem_u32(smem_a));
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      mbar_wait(bar(0), pa);
      pa ^= 1;
      tc_fence_after();
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(bar(8 + s), ph);
        const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)s * (tile_bytes >> 4));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar(5 + buf), pb ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < ksteps; ++j) {
              constexpr uint32_t kAtom16 = 16384u >> 4, kStep16 = 32u >> 4;
              const uint32_t koff16 = (uint32_t)(j >> 2) * kAtom16 + (uint32_t)(j & 3) * kStep16;
              if (TS)
                tc_mma_f16_ts(d_tmem, tmem_base + a_col0 + (uint32_t)(h * (kp >> 1) + 8 * j), b_desc + koff16,
                              kIdesc, j > 0);
              else
                tc_mma_f16(d_tmem, a_desc0 + (uint64_t)(h * (tile_bytes >> 4)) + koff16, b_desc + koff16, kIdesc,
                           j > 0);
            }
            tc_commit(bar(2 + buf));         // this half-tile's accumulators are ready
            if (h == 1) tc_commit(bar(16 + s));   // B stage free once these MMAs have read it
          }
          __syncwarp();
          if (++buf == 3) { buf = 0; pb ^= 1; }
        }
        if (++s == n_stages) { s = 0; ph ^= 1; }
      }
      if (elect_one_sync()) tc_commit(bar(1));   // A free for the next user block
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w may touch TMEM lanes 32*(w%4)..+31; warps 2-5 half 0, 6-9 half 1
    const int q = warp & 3, h = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t i = (uint32_t)h, pa = 0;          // half-tile counter of this half: 2 * tiles + h
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      const int user = blk * kUserBlock + h * 128 + q * 32 + lane;   // < u_pad by construction
      if (TS) {
        // this thread's fp16 row -> TMEM (A operand): column c holds elements 2c, 2c+1
        mbar_wait(bar(1), pa ^ 1);             // MMAs of the previous block no longer read A
        tc_fence_after();
        const uint4* src = reinterpret_cast<const uint4*>(a16 + (size_t)user * kp);
        for (int c = 0; c < (kp >> 5); ++c) {
          uint32_t w[16];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const uint4 v = __ldg(src + 4 * c + x);
            w[4 * x] = v.x; w[4 * x + 1] = v.y; w[4 * x + 2] = v.z; w[4 * x + 3] = v.w;
          }
          tc_st16(tmem_base + lane_base + a_col0 + (uint32_t)(h * (kp >> 1) + 16 * c), w);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(0));
        pa ^= 1;
      }
      const float an = anorm[user] * kEpsRel;
      for (int t = 0; t < n_tiles; ++t, i += 2) {
        const uint32_t buf = i % 3u, par = (i / 3u) & 1u;
        mbar_wait(bar(2 + buf), par);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + buf * 128u;
        const int item0 = t * kTileN;
        const bool ragged = item0 + kTileN > n_items;
        float gm[8];
        if (dbg_mode == 1) {                 // timing experiment: MMA/TMA pipeline alone
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(5 + buf));
          continue;
        }
        // 128 columns as four 32-column loads, two in flight per wait: the second pair is
        // requested before the first is reduced, so TMEM latency overlaps the FMNMX work
        uint32_t va[32], vb[32], vc[32], vd[32];
        tc_ld32(taddr, va);
        tc_ld32(taddr + 32, vb);
        const float e = fmaf(an, __ldg(btile + t), eps_abs);
        const int nv = ragged ? n_items - item0 : kTileN;     // valid columns of this tile
        tc_wait_ld64(va, vb);
        tc_ld32(taddr + 64, vc);
        tc_ld32(taddr + 96, vd);
        gm[0] = group_max16(va, nv);
        gm[1] = group_max16(va + 16, nv - 16);
        gm[2] = group_max16(vb, nv - 32);
        gm[3] = group_max16(vb + 16, nv - 48);
        tc_wait_ld64(vc, vd);
        // all TMEM reads of this buffer are complete: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(5 + buf));
        gm[4] = group_max16(vc, nv - 64);
        gm[5] = group_max16(vc + 16, nv - 80);
        gm[6] = group_max16(vd, nv - 96);
        gm[7] = group_max16(vd + 16, nv - 112);

        const float tm = fmaxf(max3(gm[0], gm[1], gm[2]), max3(max3(gm[3], gm[4], gm[5]), gm[6], gm[7]));
        if (dbg_mode == 2 && tm != 12345.678f) continue;   // timing experiment: no stores
        uint4 rec;
        {
          uint32_t* rw = reinterpret_cast<uint32_t*>(&rec);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const __half2 hh = __halves2half2(__float2half_ru((gm[2 * p] + e) * out_scale),
                                              __float2half_ru((gm[2 * p + 1] + e) * out_scale));
            rw[p] = *reinterpret_cast<const uint32_t*>(&hh);
          }
        }
        __stcs(reinterpret_cast<uint4*>(gmax16) + (size_t)t * u_pad + user, rec);
        {
          const __half2 hh = __halves2half2(__float2half_rd((tm - e) * out_scale),
                                            __float2half_ru((tm + e) * out_scale));
          gtile[(size_t)t * u_pad + user] = *reinterpret_cast<const uint32_t*>(&hh);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// max over the 128 item rows of a tile of their (scaled) norms
__global__ void k_tile_norm(const float* __restrict__ bnorm_item, int n_tiles, float* __restrict__ btile) {
  const int t = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (t >= n_tiles) return;
  float m = 0.f;
  for (int r = lane; r < kTileN; r += 32) m = fmaxf(m, bnorm_item[(size_t)t * kTileN + r]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) btile[t] = m;
}

// ------------------------------------------------------------------------------------ threshold
__device__ __forceinline__ unsigned key_of(unsigned short bits) {   // monotone fp16 -> uint16
  return (bits & 0x8000u) ? (unsigned)(~bits & 0xFFFFu) : (unsigned)(bits | 0x8000u);
}
__device__ __forceinline__ unsigned short bits_of(unsigned key) {
  return (unsigned short)((key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu));
}

// CTA = 32 users; the lower bounds of their tile maxima are staged in shared memory as monotone
// 16-bit keys, one row per user (row pitch = odd number of words: conflict-free both ways); each
// warp then selects, for 4 users, the r-th largest (r = k + n_seen) by bisection on the key bits,
// two keys per 32-bit word. T'_u = that value: at least r tiles hold an item whose true score is
// >= T'_u, at most n_seen of them through a seen item, so the k-th best MASKED score is >= T'_u.
// Small item sets (few tiles) select among the 16-item groups instead: their stored upper bounds
// are turned back into lower bounds (U - 2^-10 |U| - 2 e). Very large item sets (more tiles than keys
// fit in shared memory: > 409 600 items) select among the maxima of `fold` consecutive tiles: every
// such maximum still stands for a distinct item, so the r-th largest is a valid (slightly lower) T'_u.
__global__ void __launch_bounds__(256)
k_threshold(const uint32_t* __restrict__ gtile, const uint32_t* __restrict__ gmax16, int use_groups, int n_tiles,
            int fold, int pitch, int u_pad, int n_users, int64_t user0, int k, const int64_t* __restrict__ seen_ptr,
            const float* __restrict__ anorm, const float* __restrict__ btile, ScoreScalars* sc, float out_scale,
            float eps_abs, float* __restrict__ thr_grp, float* __restrict__ thr_exact, uint8_t* __restrict__ flag,
            int32_t* __restrict__ fb_users) {
  extern __shared__ unsigned short s_keys[];   // [32][pitch]
  const int u0 = blockIdx.x * 32;
  const int n_sel = use_groups ? n_tiles * 8 : (n_tiles + fold - 1) / fold;
  if (fold > 1) {
    for (int i = threadIdx.x; i < pitch * 32; i += 256) {
      const int t = i >> 5, j = i & 31;
      unsigned key = 0;                              // padding: below every real key
      if (t < n_sel) {
        const int t_end = min(n_tiles, (t + 1) * fold);
        for (int tt = t * fold; tt < t_end; ++tt)
          key = max(key, key_of((unsigned short)(gtile[(size_t)tt * u_pad + u0 + j] & 0xFFFFu)));
        if (key == 0) key = 1;
      }
      s_keys[j * pitch + t] = (unsigned short)key;
    }
  }
  for (int i0 = threadIdx.x; fold == 1 && i0 < pitch * 32; i0 += 1024) {
    uint32_t raw[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {                 // four loads in flight per thread
      const int i = i0 + 256 * x, t = i >> 5, j = i & 31;
      raw[x] = 0;
      if (t < n_sel)
        raw[x] = use_groups ? gmax16[((size_t)(t >> 3) * u_pad + u0 + j) * 4 + ((t >> 1) & 3)]
                            : gtile[(size_t)t * u_pad + u0 + j];
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int i = i0 + 256 * x, t = i >> 5, j = i & 31;
      if (i >= pitch * 32) continue;
      unsigned short key = 0;                    // padding: below every real key
      if (t < n_sel) {
        if (use_groups) {
          const uint32_t w = raw[x];
          const float up = __half2float(__ushort_as_half((unsigned short)((t & 1) ? (w >> 16) : (w & 0xFFFFu))));
          const float e = fmaf(anorm[u0 + j] * kEpsRel, btile[t >> 3], eps_abs) * out_scale;
          const float lo = up - 0.0009765625f * fabsf(up) - 2.f * e;
          key = (unsigned short)key_of(__half_as_ushort(__float2half_rd(lo)));
        } else {
          key = (unsigned short)key_of((unsigned short)(raw[x] & 0xFFFFu));
        }
        if (key == 0) key = 1;
      }
      s_keys[j * pitch + t] = key;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ea = scale_exp(sc->amax_bits), eb = scale_exp(sc->bmax_bits);
  const int n_words = pitch >> 1;
  for (int j = warp; j < 32; j += 8) {
    const int u = u0 + j;
    if (u >= n_users) break;
    const int n_seen = seen_ptr ? (int)(seen_ptr[user0 + u + 1] - seen_ptr[user0 + u]) : 0;
    const int r = k + n_seen;
    if (r > n_sel) {                        // the bound cannot be proven: exhaustive path
      if (lane == 0) {
        flag[u] = 1;
        thr_grp[u] = INFINITY;
        thr_exact[u] = INFINITY;
        fb_users[atomicAdd(&sc->fb_count, 1)] = u;
      }
      continue;
    }
    const unsigned short* row = s_keys + j * pitch;
    unsigned prefix = 0;
    if (pitch <= 512) {                       // keys of this user in registers: 16 per lane
      unsigned kreg[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int t = lane + 32 * i;
        kreg[i] = t < pitch ? row[t] : 0u;
      }
      for (int bit = 15; bit >= 0; --bit) {
        const unsigned cand = prefix | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) cnt += kreg[i] >= cand ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= r) prefix = cand;
      }
    } else {
      const uint32_t* row32 = reinterpret_cast<const uint32_t*>(row);
      for (int bit = 15; bit >= 0; --bit) {
        const unsigned cand = prefix | (1u << bit);
        int cnt = 0;
        for (int w = lane; w < n_words; w += 32) {
          const uint32_t v = row32[w];
          cnt += ((v & 0xFFFFu) >= cand ? 1 : 0) + ((v >> 16) >= cand ? 1 : 0);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= r) prefix = cand;
      }
    }
    if (lane == 0) {
      const float T = __half2float(__ushort_as_half(bits_of(prefix)));      // stored units, lower bound
      flag[u] = 0;
      thr_grp[u] = T;
      // back to true units: stored = score * 2^(ea+eb) * out_scale
      thr_exact[u] = scalbnf(T / out_scale, -(ea + eb));
    }
  }
}

// ------------------------------------------------------------------------------------ exact re-scoring
// Fixed summation order shared by k_rescore and k_exhaustive (identical scores): even float4
// chunks into one accumulator, odd chunks into another, x/y/z/w in order, total = even + odd.
__device__ __forceinline__ float4 load_row_f4(const float* __restrict__ row, int c, int d) {
  if (4 * c + 3 < d) return __ldg(reinterpret_cast<const float4*>(row + 4 * c));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (4 * c < d) v.x = row[4 * c];
  if (4 * c + 1 < d) v.y = row[4 * c + 1];
  if (4 * c + 2 < d) v.z = row[4 * c + 2];
  return v;
}


// Candidate (user, group) pairs bucketed by 16-item group, in two streaming passes over the
// stored bounds: FILL == false counts the hits of every group, FILL == true writes the user index
// into the group's slice of `list` (slices from the prefix sum of the counts). A tile is looked at
// in detail (its four group words) only when its own upper bound reaches the user's threshold.
// Entries that do not fit `list_cap` are dropped and their user is sent to the exhaustive path.
// Four users per thread are in flight at once (the loop is latency-bound otherwise).
#ifndef LGC_SCAN_UPT
#define LGC_SCAN_UPT 4
#endif
constexpr int kScanUPT = LGC_SCAN_UPT;       // users per thread in flight
template <bool FILL>
__global__ void __launch_bounds__(256)
k_scan(const uint32_t* __restrict__ gtile, const uint32_t* __restrict__ gmax16, int u_pad, int n_users,
       const float* __restrict__ thr_grp, int* __restrict__ grp_cnt, const int* __restrict__ grp_off,
       int* __restrict__ grp_cur, int* __restrict__ list, int list_cap, uint8_t* __restrict__ flag,
       uint8_t* __restrict__ hitmask) {
  const int tile = blockIdx.x;
  const int per = round_up((n_users + gridDim.y - 1) / gridDim.y, 256 * kScanUPT);
  const int u_beg = blockIdx.y * per, u_end = min(n_users, u_beg + per);
  __shared__ int s_cnt[8];
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int my[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (FILL) {
    // second pass: the count pass left one byte per (tile, user) with the hit groups
    for (int u0 = u_beg; u0 < u_end; u0 += 256 * kScanUPT) {
      unsigned hm[kScanUPT];
#pragma unroll
      for (int j = 0; j < kScanUPT; ++j) {
        const int u = u0 + threadIdx.x + 256 * j;
        hm[j] = u < u_end ? (unsigned)__ldcs(hitmask + (size_t)tile * u_pad + u) : 0u;
      }
#pragma unroll
      for (int j = 0; j < kScanUPT; ++j) {
        const int u = u0 + threadIdx.x + 256 * j;
        unsigned hits = hm[j];
        while (hits) {
          const int g = __ffs(hits) - 1;
          hits &= hits - 1;
          const int pos = grp_off[tile * 8 + g] + atomicAdd(&grp_cur[tile * 8 + g], 1);
          if (pos < list_cap) list[pos] = u;
          else flag[u] = 2;                            // picked up by k_select -> exhaustive path
        }
      }
    }
    return;
  }
  for (int u0 = u_beg; u0 < u_end; u0 += 256 * kScanUPT) {
    float thr[kScanUPT]; uint32_t tw[kScanUPT];
#pragma unroll
    for (int j = 0; j < kScanUPT; ++j) {
      const int u = u0 + threadIdx.x + 256 * j;
      thr[j] = INFINITY; tw[j] = 0xFC00FC00u;          // -inf bounds
      if (u < u_end) { thr[j] = thr_grp[u]; tw[j] = __ldcs(gtile + (size_t)tile * u_pad + u); }
    }
#pragma unroll
    for (int j = 0; j < kScanUPT; ++j) {
      const int u = u0 + threadIdx.x + 256 * j;
      const float up = __half2float(__ushort_as_half((unsigned short)(tw[j] >> 16)));
      unsigned hits = 0;
      if (up >= thr[j]) {
        const uint4 rec = __ldcs(reinterpret_cast<const uint4*>(gmax16) + (size_t)tile * u_pad + u);
        const uint32_t w[4] = {rec.x, rec.y, rec.z, rec.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[p]));
          if (f.x >= thr[j]) hits |= 1u << (2 * p);
          if (f.y >= thr[j]) hits |= 1u << (2 * p + 1);
        }
      }
      if (u < u_end) hitmask[(size_t)tile * u_pad + u] = (uint8_t)hits;
#pragma unroll
      for (int g = 0; g < 8; ++g) my[g] += (hits >> g) & 1;
    }
  }
  if (!FILL) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int v = __reduce_add_sync(0xffffffffu, my[g]);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[g], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&grp_cnt[tile * 8 + threadIdx.x], s_cnt[threadIdx.x]);
  }
}

// exclusive prefix sum of the per-group counts (a few thousand values)
__global__ void k_group_prefix(const int* __restrict__ grp_cnt, int n_groups, int* __restrict__ grp_off,
                               ScoreScalars* __restrict__ sc) {
  __shared__ long long s_part[256];
  const int per = (n_groups + 255) / 256;
  const int b = min(n_groups, (int)threadIdx.x * per), e = min(n_groups, b + per);
  long long sum = 0;
  for (int t = b; t < e; ++t) sum += grp_cnt[t];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < 256; ++i) { const long long v = s_part[i]; s_part[i] = run; run += v; }
    sc->n_groups += (unsigned long long)run;
    grp_off[n_groups] = (int)min(run, (long long)0x7fffffff);
  }
  __syncthreads();
  long long run = s_part[threadIdx.x];
  for (int t = b; t < e; ++t) {
    grp_off[t] = (int)min(run, (long long)0x7fffffff);
    run += grp_cnt[t];
  }
}

constexpr int kRowsPerWarp = 16;      // user rows a warp keeps in flight

// grid (n_tiles, splits). The CTA keeps its 128 item rows (fp32) in shared memory and walks the
// candidate lists of the tile's 8 groups. Per step a warp takes 16 users of ONE group: their
// metadata, then all 16 user rows, are requested before anything is used (that hides the HBM
// latency of these random 256-byte reads); lane = (item of the group, half-warp) then streams its
// item row once and reuses every chunk for the 8 user rows of its half-warp -- shared-memory
// traffic is what bounds this kernel, not the 64 FMAs per score.
struct RescoreMeta {          // lanes 0..15 own one user of the step each
  int u; int64_t uid, sb; float thr; int n_seen;
};

#ifndef LGC_RESCORE_OCC
#define LGC_RESCORE_OCC 2
#endif
template <bool PIPE>
__global__ void __launch_bounds__(256, LGC_RESCORE_OCC)
k_rescore(const int* __restrict__ list, const int* __restrict__ grp_off, int n_tiles, int list_cap, int64_t user0,
          int n_items,
          int d, const float* __restrict__ user_emb, int ld_user, const int64_t* __restrict__ user_ids,
          const float* __restrict__ item_emb, int ld_item, const float* __restrict__ thr_exact,
          const int64_t* __restrict__ seen_ptr, const int64_t* __restrict__ seen_items,
          const uint8_t* __restrict__ flag, int* __restrict__ cand_cnt, int32_t* __restrict__ cand_item,
          float* __restrict__ cand_score, ScoreScalars* __restrict__ sc) {
  extern __shared__ __align__(16) uint8_t smem_rs[];
  const int d4 = (d + 3) >> 2;
  const int row_f4 = d4 + 1;                          // +16 B per row: conflict-free float4 reads
  float4* s_items = reinterpret_cast<float4*>(smem_rs);                 // [128][row_f4]
  float4* s_rows = s_items + 128 * row_f4;                              // [8 warps][16][d4]

  // Load balance: candidate lists are heavily skewed towards the tiles of popular items (a trained
  // model sends most users to the same few hundred items), so CTAs do not own tiles: CTA c owns
  // the slice [c * Q, (c + 1) * Q) of the GLOBAL list (sorted by group, hence by tile) and walks
  // the tiles its slice touches, re-staging the 128 item rows when the tile changes.
  const int n_groups = n_tiles * 8;
  const long long total = min((long long)grp_off[n_groups], (long long)list_cap);
  long long q = (total + gridDim.x - 1) / gridDim.x;
  q = (q + kRowsPerWarp - 1) / kRowsPerWarp * kRowsPerWarp;
  const long long lo_ll = (long long)blockIdx.x * q, hi_ll = min(total, lo_ll + q);
  if (lo_ll >= hi_ll) return;
  const int lo = (int)lo_ll, hi = (int)hi_ll;
  // first group whose slice reaches beyond `lo`
  int g_lo = 0;
  {
    int a = 0, b = n_groups;
    while (a < b) { const int m = (a + b) >> 1; if (grp_off[m + 1] <= lo) a = m + 1; else b = m; }
    g_lo = a;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = lane & 15, hw = lane >> 4;
  float4* rows = s_rows + (size_t)warp * kRowsPerWarp * d4;
  unsigned long long my_emit = 0;
  const int step = 8 * kRowsPerWarp;
  auto load_meta = [&](int e0, int end) {
    RescoreMeta m; m.u = 0; m.uid = 0; m.sb = 0; m.thr = INFINITY; m.n_seen = 0;
    if (e0 < end && lane < min(kRowsPerWarp, end - e0)) {
      m.u = list[e0 + lane];
      if (flag[m.u] == 0) {
        m.uid = user_ids ? user_ids[user0 + m.u] : user0 + m.u;
        m.thr = thr_exact[m.u];
        if (seen_ptr) { m.sb = seen_ptr[user0 + m.u]; m.n_seen = (int)(seen_ptr[user0 + m.u + 1] - m.sb); }
      }
    }
    return m;
  };
  // half-warp hw requests float4 column c of rows hw, hw+2, ... (8 loads in flight per lane)
  auto request_rows = [&](const RescoreMeta& m, int ne, int c, float4 (&v)[kRowsPerWarp / 2]) {
#pragma unroll
    for (int i = 0; i < kRowsPerWarp / 2; ++i) {
      const int slot = 2 * i + hw;
      const int64_t id = __shfl_sync(0xffffffffu, m.uid, slot);
      if (slot < ne && c < d4) v[i] = load_row_f4(user_emb + (size_t)id * ld_user, c, d);
    }
  };
  auto store_rows = [&](int ne, int c, const float4 (&v)[kRowsPerWarp / 2]) {
#pragma unroll
    for (int i = 0; i < kRowsPerWarp / 2; ++i) {
      const int slot = 2 * i + hw;
      if (slot < ne && c < d4) rows[slot * d4 + c] = v[i];
    }
  };

#pragma unroll 1
  for (int tile = g_lo >> 3; tile < n_tiles && min(grp_off[tile * 8], list_cap) < hi; ++tile) {
  const int item0 = tile * kTileN;
  __syncthreads();                                   // every warp is done with the previous tile's rows
  for (int i = threadIdx.x; i < 128 * d4; i += 256) {
    const int r = i / d4, c = i % d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (item0 + r < n_items) v = load_row_f4(item_emb + (size_t)(item0 + r) * ld_item, c, d);
    s_items[r * row_f4 + c] = v;
  }
  __syncthreads();
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    const int beg = max(lo, min(grp_off[tile * 8 + g], list_cap));
    const int end = min(hi, min(grp_off[tile * 8 + g + 1], list_cap));
    const int local = g * kGroup + it, item = item0 + local;
    const float4* irow = s_items + local * row_f4;
    int e0 = beg + warp * kRowsPerWarp;
    if (e0 >= end) continue;
    // software pipeline (PIPE: d <= 64, one float4 column per lane and row): metadata two steps
    // ahead, user rows one step ahead, so neither latency is exposed behind the dot products
    RescoreMeta m0 = load_meta(e0, end), m1 = load_meta(e0 + step, end);
    float4 vreg[kRowsPerWarp / 2];
    if (PIPE) request_rows(m0, min(kRowsPerWarp, end - e0), it, vreg);
    for (; e0 < end; e0 += step) {
      const int ne = min(kRowsPerWarp, end - e0);
      __syncwarp();
      if (PIPE) {
        store_rows(ne, it, vreg);
      } else {
        for (int c0 = 0; c0 < d4; c0 += 16) {
          request_rows(m0, ne, c0 + it, vreg);
          store_rows(ne, c0 + it, vreg);
        }
      }
      __syncwarp();
      const int e1 = e0 + step;
      if (PIPE && e1 < end) request_rows(m1, min(kRowsPerWarp, end - e1), it, vreg);
      const RescoreMeta m2 = load_meta(e1 + step, end);
      // 8 dot products per lane; even chunks -> s0, odd chunks -> s1 (the order of k_exhaustive)
      float s0[kRowsPerWarp / 2], s1[kRowsPerWarp / 2];
#pragma unroll
      for (int i = 0; i < kRowsPerWarp / 2; ++i) s0[i] = s1[i] = 0.f;
      int c = 0;
      for (; c + 1 < d4; c += 2) {
        const float4 y0 = irow[c], y1 = irow[c + 1];
#pragma unroll
        for (int i = 0; i < kRowsPerWarp / 2; ++i) {
          const float4 x0 = rows[(2 * i + hw) * d4 + c], x1 = rows[(2 * i + hw) * d4 + c + 1];
          s0[i] = fmaf(x0.x, y0.x, s0[i]); s0[i] = fmaf(x0.y, y0.y, s0[i]);
          s0[i] = fmaf(x0.z, y0.z, s0[i]); s0[i] = fmaf(x0.w, y0.w, s0[i]);
          s1[i] = fmaf(x1.x, y1.x, s1[i]); s1[i] = fmaf(x1.y, y1.y, s1[i]);
          s1[i] = fmaf(x1.z, y1.z, s1[i]); s1[i] = fmaf(x1.w, y1.w, s1[i]);
        }
      }
      if (c < d4) {
        const float4 y0 = irow[c];
#pragma unroll
        for (int i = 0; i < kRowsPerWarp / 2; ++i) {
          const float4 x0 = rows[(2 * i + hw) * d4 + c];
          s0[i] = fmaf(x0.x, y0.x, s0[i]); s0[i] = fmaf(x0.y, y0.y, s0[i]);
          s0[i] = fmaf(x0.z, y0.z, s0[i]); s0[i] = fmaf(x0.w, y0.w, s0[i]);
        }
      }
      // emission in two phases so the 8 slot-allocating atomics of a lane overlap
      int e_u[kRowsPerWarp / 2], pos[kRowsPerWarp / 2];
      float sco[kRowsPerWarp / 2];
#pragma unroll
      for (int i = 0; i < kRowsPerWarp / 2; ++i) {
        const int slot = 2 * i + hw;
        e_u[i] = __shfl_sync(0xffffffffu, m0.u, slot);
        const float e_thr = __shfl_sync(0xffffffffu, m0.thr, slot);
        const int e_ns = __shfl_sync(0xffffffffu, m0.n_seen, slot);
        sco[i] = s0[i] + s1[i];
        bool hit = slot < ne && item < n_items && sco[i] >= e_thr;
        if (__any_sync(0xffffffffu, hit && e_ns > 0)) {          // seen lists are short and rare
          const int64_t e_sb = __shfl_sync(0xffffffffu, m0.sb, slot);
          if (hit)
            for (int q = 0; q < e_ns; ++q) hit &= (seen_items[e_sb + q] != item);
        }
        pos[i] = hit ? atomicAdd(&cand_cnt[e_u[i]], 1) : kCap;
      }
#pragma unroll
      for (int i = 0; i < kRowsPerWarp / 2; ++i) {
        if (pos[i] < kCap) {
          cand_item[(size_t)e_u[i] * kCap + pos[i]] = item;
          cand_score[(size_t)e_u[i] * kCap + pos[i]] = sco[i] + 0.f;      // -0.0 -> +0.0
          ++my_emit;
        }
      }
      m0 = m1; m1 = m2;
    }
  }
  }   // tiles of this CTA's slice
  if (my_emit) atomicAdd(&sc->n_emitted, my_emit);
}

// ------------------------------------------------------------------------------------ final selection
// Strict total order: a before b iff (a.s > b.s) or (a.s == b.s and a.item < b.item).
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}
// (score, item) -> 64-bit key whose DEscending order is that total order; 0 = empty slot
__device__ __forceinline__ unsigned long long sel_key(float s, int item) {
  const unsigned b = __float_as_uint(s + 0.f);
  const unsigned ord = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
  return ((unsigned long long)ord << 32) | (unsigned)(0x7FFFFFFF - item);
}
__device__ __forceinline__ float sel_score(unsigned long long key) {
  const unsigned ord = (unsigned)(key >> 32);
  return __uint_as_float((ord & 0x80000000u) ? (ord ^ 0x80000000u) : ~ord);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
  const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}

// One warp per user. Entries = exact candidates + the user's seen items with masked score 0.0
// (reference: pred * (1 - mask), src/lightgcn.py:175); rows of seen_items hold distinct ids.
// Up to 64 entries: bitonic sort across the warp (1 or 2 keys per lane). More (long seen lists):
// k rounds, each picking the first entry strictly after the previous pick.
__global__ void __launch_bounds__(256)
k_select(int n_users, int64_t user0, int n_items, int k, uint8_t* flag, const int* __restrict__ cand_cnt,
         const int32_t* __restrict__ cand_item, const float* __restrict__ cand_score,
         const int64_t* __restrict__ seen_ptr, const int64_t* __restrict__ seen_items,
         int64_t* __restrict__ topk_items, float* __restrict__ topk_scores, int32_t* __restrict__ fb_users,
         ScoreScalars* __restrict__ sc) {
  const int u = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (u >= n_users || flag[u] == 1) return;          // 1: already on the exhaustive list
  const int cnt = cand_cnt[u];
  if (cnt > kCap || flag[u] == 2) {                   // 2: candidate entries were dropped
    if (lane == 0) { flag[u] = 1; fb_users[atomicAdd(&sc->fb_count, 1)] = u; }
    return;
  }
  const int64_t sb = seen_ptr ? seen_ptr[user0 + u] : 0, se = seen_ptr ? seen_ptr[user0 + u + 1] : 0;
  const int n_seen = (int)(se - sb), total = cnt + n_seen;
  if (total <= 64) {
    auto entry = [&](int e) -> unsigned long long {
      if (e < cnt) return sel_key(cand_score[(size_t)u * kCap + e], cand_item[(size_t)u * kCap + e]);
      if (e < total) {
        const int64_t si = seen_items[sb + (e - cnt)];
        return (si >= 0 && si < n_items) ? sel_key(0.f, (int)si) : 0ull;
      }
      return 0ull;
    };
    unsigned long long k0 = entry(lane), k1 = total > 32 ? entry(lane + 32) : 0ull;
    const int top = total > 32 ? 64 : 32;
    for (int size = 2; size <= top; size <<= 1) {
      for (int stride = size >> 1; stride; stride >>= 1) {
        if (stride == 32) {                       // partner is this lane's other key
          const unsigned long long hi = max(k0, k1), lo = min(k0, k1);
          k0 = hi; k1 = lo;
          continue;
        }
        const bool lower = (lane & stride) == 0;
        {
          const unsigned long long o = shfl_xor_u64(k0, stride);
          const bool desc = (lane & size) == 0;     // index x = lane
          k0 = (lower == desc) ? max(k0, o) : min(k0, o);
        }
        if (top == 64) {
          const unsigned long long o = shfl_xor_u64(k1, stride);
          const bool desc = ((lane + 32) & size) == 0;   // index x = lane + 32
          k1 = (lower == desc) ? max(k1, o) : min(k1, o);
        }
      }
    }
    if (lane < k) {
      const bool found = k0 != 0ull;
      topk_items[(size_t)(user0 + u) * k + lane] = found ? (int64_t)(0x7FFFFFFF - (int)(unsigned)k0) : -1;
      if (topk_scores) topk_scores[(size_t)(user0 + u) * k + lane] = found ? sel_score(k0) : -INFINITY;
    }
    return;
  }
  float ps = INFINITY; int pi = -1;
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY; int bi = 0x7fffffff;
    for (int e = lane; e < total; e += 32) {
      float s; int itm;
      if (e < cnt) { s = cand_score[(size_t)u * kCap + e]; itm = cand_item[(size_t)u * kCap + e]; }
      else {
        const int64_t si = seen_items[sb + (e - cnt)];
        if (si < 0 || si >= n_items) continue;
        s = 0.f; itm = (int)si;
      }
      if (before(ps, pi, s, itm) && before(s, itm, bs, bi)) { bs = s; bi = itm; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (before(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    const bool found = bi != 0x7fffffff;
    if (lane == 0) {
      topk_items[(size_t)(user0 + u) * k + r] = found ? bi : -1;
      if (topk_scores) topk_scores[(size_t)(user0 + u) * k + r] = found ? bs : -INFINITY;
    }
    ps = bs; pi = bi;
  }
}

// ------------------------------------------------------------------------------------ exhaustive path
// Persistent CTAs over the fallback list: all masked fp32 scores of one user into a scratch row,
// then k rounds of block-wide selection in the same total order.
__global__ void __launch_bounds__(1024)
k_exhaustive(const int32_t* __restrict__ fb_users, const ScoreScalars* __restrict__ sc, int64_t user0,
             int n_items, int d, int k, const float* __restrict__ user_emb, int ld_user,
             const int64_t* __restrict__ user_ids, const float* __restrict__ item_emb, int ld_item,
             const int64_t* __restrict__ seen_ptr, const int64_t* __restrict__ seen_items,
             float* __restrict__ scratch, int64_t* __restrict__ topk_items, float* __restrict__ topk_scores) {
  extern __shared__ __align__(16) uint8_t smem_ex[];
  const int d4 = (d + 3) >> 2;
  float4* s_user = reinterpret_cast<float4*>(smem_ex);
  __shared__ float s_bs[32];
  __shared__ int s_bi[32];
  __shared__ float s_ps;
  __shared__ int s_pi;
  float* my = scratch + (size_t)blockIdx.x * n_items;
  const int n_fb = sc->fb_count;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec_ok = (ld_item % 4 == 0) && ((uintptr_t)item_emb % 16 == 0) && (d % 4 == 0);
  for (int f = blockIdx.x; f < n_fb; f += gridDim.x) {
    const int u = fb_users[f];
    const int64_t uid = user_ids ? user_ids[user0 + u] : user0 + u;
    const float* urow = user_emb + (size_t)uid * ld_user;
    __syncthreads();
    for (int c = threadIdx.x; c < d4; c += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = urow[4 * c];
      if (4 * c + 1 < d) v.y = urow[4 * c + 1];
      if (4 * c + 2 < d) v.z = urow[4 * c + 2];
      if (4 * c + 3 < d) v.w = urow[4 * c + 3];
      s_user[c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
      const float* irow = item_emb + (size_t)i * ld_item;
      float s0 = 0.f, s1 = 0.f;
      for (int c = 0; c < d4; ++c) {
        float4 y;
        if (vec_ok) y = __ldg(reinterpret_cast<const float4*>(irow) + c);
        else {
          y = make_float4(0.f, 0.f, 0.f, 0.f);
          y.x = irow[4 * c];
          if (4 * c + 1 < d) y.y = irow[4 * c + 1];
          if (4 * c + 2 < d) y.z = irow[4 * c + 2];
          if (4 * c + 3 < d) y.w = irow[4 * c + 3];
        }
        const float4 x = s_user[c];
        float& s = (c & 1) ? s1 : s0;
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
      }
      my[i] = s0 + s1;
    }
    __syncthreads();
    if (seen_ptr) {
      for (int64_t q = seen_ptr[user0 + u] + threadIdx.x; q < seen_ptr[user0 + u + 1]; q += blockDim.x) {
        const int64_t si = seen_items[q];
        if (si >= 0 && si < n_items) my[si] = 0.f;          // multiplicative mask
      }
    }
    if (threadIdx.x == 0) { s_ps = INFINITY; s_pi = -1; }
    __syncthreads();
    for (int r = 0; r < k; ++r) {
      const float ps = s_ps; const int pi = s_pi;
      float bs = -INFINITY; int bi = 0x7fffffff;
      for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
        const float s = my[i];
        if (before(ps, pi, s, i) && before(s, i, bs, bi)) { bs = s; bi = i; }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (before(os, oi, bs, bi)) { bs = os; bi = oi; }
      }
      if (lane == 0) { s_bs[warp] = bs; s_bi[warp] = bi; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
          if (before(s_bs[w], s_bi[w], bs, bi)) { bs = s_bs[w]; bi = s_bi[w]; }
        const bool found = bi != 0x7fffffff;
        topk_items[(size_t)(user0 + u) * k + r] = found ? bi : -1;
        if (topk_scores) topk_scores[(size_t)(user0 + u) * k + r] = found ? bs : -INFINITY;
        s_ps = bs; s_pi = bi;
      }
      __syncthreads();
    }
  }
}

__global__ void k_add_stats(const ScoreScalars* __restrict__ sc, int64_t* __restrict__ stats) {
  stats[0] += sc->fb_count;
  stats[1] += (int64_t)sc->n_groups;
  stats[2] += (int64_t)sc->n_emitted;
  stats[3] += 1;
}

// ------------------------------------------------------------------------------------ host side
struct Layout {
  int kp, katoms, n_stages, n_tiles, i_pad, chunk, chunk_pad, ts, list_cap;
  size_t gemm_smem;
  size_t off_scal, off_b16, off_bnorm, off_btile, off_a16, off_anorm, off_g16, off_g128, off_tcnt, off_toff, off_tcur, off_list, off_hit, off_thr_grp, off_thr_exact, off_flag,
      off_fb, off_cnt, off_citem, off_cscore, off_scratch;
  int n_exh_ctas;
  size_t bytes;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

bool make_layout(int64_t n_users, int64_t n_items, int d, int k, Layout* L) {
  if (n_users <= 0 || n_items <= 0 || d <= 0 || d > 256 || k < 1 || k > 32 || k > n_items) return false;
  if (n_items > (1LL << 28) || n_users > (1LL << 40)) return false;
  L->kp = round_up(d, 64);
  L->katoms = L->kp / 64;
  const size_t tile_bytes = 16384u * L->katoms;
  const size_t budget = 200 * 1024;
  L->ts = L->kp <= 128 ? 1 : 0;          // A operand in TMEM (fits beside three accumulators)
  const size_t a_tiles = L->ts ? 0 : 2;
  int ns = (int)((budget - a_tiles * tile_bytes) / tile_bytes);
  L->n_stages = std::max(1, std::min(ns, 8));
  L->gemm_smem = 1024 + (a_tiles + (size_t)L->n_stages) * tile_bytes + 256;
  L->i_pad = round_up(n_items, kTileN);
  L->n_tiles = L->i_pad / kTileN;
  // bytes per user of the chunk-sized buffers
  const size_t per_user = (size_t)L->kp * 2 + 4 + (size_t)L->n_tiles * 4 * 4 + (size_t)L->n_tiles * 4 + 4 + 4 + 1 +
                          4 + 4 + (size_t)kCap * 8 + 40 * 4 + (size_t)L->n_tiles;
  const size_t fixed = (size_t)L->i_pad * (L->kp * 2 + 8) + (size_t)296 * n_items * 4 + (1 << 16);
  int64_t chunk = round_up(n_users, kUserBlock);
  if (fixed + per_user * (size_t)chunk > kWorkspaceBudget) {
    int64_t c = (int64_t)((kWorkspaceBudget > fixed ? kWorkspaceBudget - fixed : 0) / per_user);
    c = c / (kUserBlock * 148) * (kUserBlock * 148);      // whole waves of user blocks
    chunk = std::max<int64_t>(c, kUserBlock * 148);
  }
  if (chunk > (1 << 28)) chunk = 1 << 28;
  L->chunk = (int)std::min<int64_t>(chunk, round_up(n_users, kUserBlock));
  L->chunk_pad = round_up(L->chunk, kUserBlock);
  L->n_exh_ctas = 296;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += align256(b); return o; };
  const size_t cp = (size_t)L->chunk_pad;
  L->off_scal = take(sizeof(ScoreScalars));
  L->off_b16 = take((size_t)L->i_pad * L->kp * 2);
  L->off_bnorm = take((size_t)L->i_pad * 4);
  L->off_btile = take((size_t)L->n_tiles * 4);
  L->off_a16 = take(cp * L->kp * 2);
  L->off_anorm = take(cp * 4);
  L->off_g16 = take((size_t)L->n_tiles * 4 * cp * 4);
  L->off_g128 = take((size_t)L->n_tiles * cp * 4);
  L->list_cap = (int)std::min<size_t>((size_t)0x7ffffff0, std::max<size_t>((size_t)1 << 16, cp * 40));
  L->off_tcnt = take((size_t)(L->n_tiles * 8 + 1) * 4);
  L->off_toff = take((size_t)(L->n_tiles * 8 + 1) * 4);
  L->off_tcur = take((size_t)(L->n_tiles * 8 + 1) * 4);
  L->off_list = take((size_t)L->list_cap * 4);
  L->off_hit = take((size_t)L->n_tiles * cp);
  L->off_thr_grp = take(cp * 4);
  L->off_thr_exact = take(cp * 4);
  L->off_flag = take(cp);
  L->off_fb = take(cp * 4);
  L->off_cnt = take(cp * 4);
  L->off_citem = take(cp * kCap * 4);
  L->off_cscore = take(cp * kCap * 4);
  L->off_scratch = take((size_t)L->n_exh_ctas * n_items * 4);
  L->bytes = off + 1024;
  return true;
}

int encode_map(CUtensorMap* map, const void* base, int kp, int64_t rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kp * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  // The driver entry point is resolved through the runtime, so the library has no link-time
  // dependency on libcuda.so (it must load on a machine without a GPU driver).
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LGC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available in this driver");
      return LGC_ERR_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return LGC_ERR_CUDA;
  }
  return LGC_OK;
}

}  // namespace
}  // namespace lgc

using namespace lgc;

extern "C" size_t lgc_score_topk_workspace_bytes(int64_t n_users, int64_t n_items, int d, int k) {
  Layout L;
  if (!make_layout(n_users, n_items, d, k, &L)) return 0;
  return L.bytes;
}

extern "C" int lgc_score_topk(const lgc_score_topk_args* a, void* stream) {
  NvtxRange nvtx("lgc_score_topk");
  LGC_REQUIRE(a, "null argument");
  LGC_REQUIRE(a->user_emb && a->item_emb && a->topk_items && a->workspace, "null field in lgc_score_topk_args");
  LGC_REQUIRE(a->d >= 1 && a->d <= a->ld_user && a->d <= a->ld_item, "d must fit both row strides");
  LGC_REQUIRE((a->seen_ptr == nullptr) == (a->seen_items == nullptr) || a->seen_ptr, "seen_items without seen_ptr");
  Layout L;
  if (!make_layout(a->n_users, a->n_items, a->d, a->k, &L)) {
    set_error("lgc_score_topk: need 1 <= k <= min(32, n_items), 1 <= d <= 256, n_users > 0");
    return LGC_ERR_INVALID;
  }
  if (a->workspace_bytes < L.bytes) {
    set_error("lgc_score_topk: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)(((uintptr_t)a->workspace + 1023) & ~(uintptr_t)1023);
  ScoreScalars* sc = (ScoreScalars*)(ws + L.off_scal);
  __half* b16 = (__half*)(ws + L.off_b16);
  __half* a16 = (__half*)(ws + L.off_a16);
  float* bnorm_item = (float*)(ws + L.off_bnorm);
  float* btile = (float*)(ws + L.off_btile);
  float* anorm = (float*)(ws + L.off_anorm);
  uint32_t* g16 = (uint32_t*)(ws + L.off_g16);
  uint32_t* g128 = (uint32_t*)(ws + L.off_g128);
  int* tile_cnt = (int*)(ws + L.off_tcnt);
  int* tile_off = (int*)(ws + L.off_toff);
  int* tile_cur = (int*)(ws + L.off_tcur);
  int* list = (int*)(ws + L.off_list);
  uint8_t* hitmask = ws + L.off_hit;
  float* thr_grp = (float*)(ws + L.off_thr_grp);
  float* thr_exact = (float*)(ws + L.off_thr_exact);
  uint8_t* flag = ws + L.off_flag;
  int32_t* fb = (int32_t*)(ws + L.off_fb);
  int* cnt = (int*)(ws + L.off_cnt);
  int32_t* citem = (int32_t*)(ws + L.off_citem);
  float* cscore = (float*)(ws + L.off_cscore);
  float* scratch = (float*)(ws + L.off_scratch);

  const bool f4_ok = a->ld_user % 4 == 0 && a->ld_item % 4 == 0 && (uintptr_t)a->user_emb % 16 == 0 &&
                     (uintptr_t)a->item_emb % 16 == 0;
  LGC_REQUIRE(f4_ok, "embedding rows must be 16-byte aligned (ld % 4 == 0)");

  static bool attr_done_dev[kMaxDevices] = {};       // cudaFuncSetAttribute is per device
  bool& attr_done = attr_done_dev[current_device_slot()];
  if (!attr_done) {
    LGC_CUDA(cudaFuncSetAttribute(k_score_gemm<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_score_gemm<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_score_gemm<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_score_gemm<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_threshold, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_rescore<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_rescore<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }

  CUtensorMap map_a, map_b;
  int rc = encode_map(&map_a, a16, L.kp, L.chunk_pad);
  if (rc) return rc;
  rc = encode_map(&map_b, b16, L.kp, L.i_pad);
  if (rc) return rc;

  const char* dbg_env = getenv("LGC_SCORE_DEBUG_MODE");   // timing experiments only (results invalid)
  const int dbg_mode = dbg_env ? atoi(dbg_env) : 0;
  const int out_e = out_scale_exp(L.kp);
  const float out_scale = ldexpf(1.0f, out_e);

  LGC_CUDA(cudaMemsetAsync(sc, 0, sizeof(ScoreScalars), st));
  {
    ProfScope ps(PROF_SCORE_CONVERT, st);
    k_absmax<<<kNumSMs * 4, 256, 0, st>>>(a->user_emb, a->ld_user, a->d, a->user_ids, a->n_users, &sc->amax_bits);
    LGC_LAUNCH_CHECK();
    k_absmax<<<kNumSMs * 4, 256, 0, st>>>(a->item_emb, a->ld_item, a->d, nullptr, a->n_items, &sc->bmax_bits);
    LGC_LAUNCH_CHECK();
    k_convert<<<(int)ceil_div((int64_t)L.i_pad * 32, 256), 256, 0, st>>>(
        a->item_emb, a->ld_item, a->d, nullptr, 0, a->n_items, L.i_pad, L.kp, &sc->bmax_bits, b16, bnorm_item,
        &sc->bnorm_bits);
    LGC_LAUNCH_CHECK();
    k_tile_norm<<<(int)ceil_div((int64_t)L.n_tiles * 32, 256), 256, 0, st>>>(bnorm_item, L.n_tiles, btile);
    LGC_LAUNCH_CHECK();
  }

  const int use_groups = L.n_tiles < 256 ? 1 : 0;
  // keys per user in k_threshold's shared memory: at most 3 200 (200 KB for 32 users); beyond 409 600 items
  // the threshold is selected among the maxima of `fold` consecutive tiles (LGC_SCORE_FOLD: test override)
  const char* fold_env = getenv("LGC_SCORE_FOLD");
  int fold = use_groups ? 1 : (int)ceil_div(L.n_tiles, 3200);
  if (fold_env && !use_groups) fold = std::max(fold, std::min(atoi(fold_env), std::max(1, L.n_tiles / 64)));
  const int n_sel = use_groups ? L.n_tiles * 8 : (int)ceil_div(L.n_tiles, fold);
  int pitch = (n_sel + 1) / 2 * 2;                 // keys per user row: even, with an odd word count
  if ((pitch / 2) % 2 == 0) pitch += 2;
  const size_t thr_smem = (size_t)pitch * 32 * 2;
  const float eps_abs = (float)L.kp * 0.001953125f;   // subnormal fp16 inputs: kp * 2^-9 (scaled units)
  LGC_REQUIRE(thr_smem <= 200 * 1024, "lgc_score_topk: threshold keys do not fit shared memory");
  const int d4 = (a->d + 3) / 4;
  const size_t rs_smem = (size_t)128 * (d4 + 1) * 16 + (size_t)8 * kRowsPerWarp * d4 * 16;
  const size_t ex_smem = (size_t)d4 * 16;

  for (int64_t user0 = 0; user0 < a->n_users; user0 += L.chunk) {
    const int nu = (int)std::min<int64_t>(L.chunk, a->n_users - user0);
    const int nu_pad = round_up(nu, kUserBlock);
    const int n_blocks = nu_pad / kUserBlock;
    // chunk-local state; u_pad of the stored maxima is always L.chunk_pad (fixed row pitch)
    LGC_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nu_pad * 4, st));
    LGC_CUDA(cudaMemsetAsync(&sc->fb_count, 0, sizeof(int), st));
    {
      ProfScope ps(PROF_SCORE_CONVERT, st);
      k_convert<<<(int)ceil_div((int64_t)nu_pad * 32, 256), 256, 0, st>>>(
          a->user_emb, a->ld_user, a->d, a->user_ids, user0, nu, nu_pad, L.kp, &sc->amax_bits, a16, anorm, nullptr);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_GEMM, st);
      const int grid = std::min(n_blocks, kNumSMs);
#define LGC_GEMM_LAUNCH(TS_, KA_)                                                                          \
  k_score_gemm<TS_, KA_><<<grid, kGemmThreads, L.gemm_smem, st>>>(                                         \
      map_a, map_b, a16, anorm, btile, n_blocks, L.n_tiles, (int)a->n_items, L.n_stages, L.chunk_pad,      \
      out_scale, eps_abs, g16, g128, dbg_mode)
      if (L.katoms == 1) LGC_GEMM_LAUNCH(true, 1);
      else if (L.katoms == 2) LGC_GEMM_LAUNCH(true, 2);
      else if (L.katoms == 3) LGC_GEMM_LAUNCH(false, 3);
      else LGC_GEMM_LAUNCH(false, 4);
#undef LGC_GEMM_LAUNCH
      LGC_LAUNCH_CHECK();
    }
    if (dbg_mode) continue;          // timing experiments: GEMM only, outputs are not valid
    {
      ProfScope ps(PROF_SCORE_SELECT, st);
      k_threshold<<<nu_pad / 32, 256, thr_smem, st>>>(g128, g16, use_groups, L.n_tiles, fold, pitch, L.chunk_pad, nu, user0,
                                                     a->k, a->seen_ptr, anorm, btile, sc, out_scale, eps_abs, thr_grp,
                                                     thr_exact, flag, fb);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_SCAN, st);
      LGC_CUDA(cudaMemsetAsync(tile_cnt, 0, (size_t)(L.n_tiles * 8 + 1) * 4, st));
      LGC_CUDA(cudaMemsetAsync(tile_cur, 0, (size_t)(L.n_tiles * 8 + 1) * 4, st));
      const int ssplits = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(kNumSMs * 16, L.n_tiles),
                                                                      ceil_div(nu, 1024)));
      dim3 sgrid(L.n_tiles, ssplits);
      k_scan<false><<<sgrid, 256, 0, st>>>(g128, g16, L.chunk_pad, nu, thr_grp, tile_cnt, tile_off, tile_cur, list,
                                          L.list_cap, flag, hitmask);
      LGC_LAUNCH_CHECK();
      k_group_prefix<<<1, 256, 0, st>>>(tile_cnt, L.n_tiles * 8, tile_off, sc);
      LGC_LAUNCH_CHECK();
      k_scan<true><<<sgrid, 256, 0, st>>>(g128, g16, L.chunk_pad, nu, thr_grp, tile_cnt, tile_off, tile_cur, list,
                                         L.list_cap, flag, hitmask);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_RESCORE, st);
      const int grid = kNumSMs * 8;        // slices of the global candidate list (4 waves of 2 CTAs/SM)
      if (d4 <= 16)
        k_rescore<true><<<grid, 256, rs_smem, st>>>(list, tile_off, L.n_tiles, L.list_cap, user0, (int)a->n_items,
                                                    a->d, a->user_emb, a->ld_user, a->user_ids, a->item_emb,
                                                    a->ld_item, thr_exact, a->seen_ptr, a->seen_items, flag, cnt,
                                                    citem, cscore, sc);
      else
        k_rescore<false><<<grid, 256, rs_smem, st>>>(list, tile_off, L.n_tiles, L.list_cap, user0, (int)a->n_items,
                                                     a->d, a->user_emb, a->ld_user, a->user_ids, a->item_emb,
                                                     a->ld_item, thr_exact, a->seen_ptr, a->seen_items, flag, cnt,
                                                     citem, cscore, sc);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_FINAL, st);
      k_select<<<(int)ceil_div((int64_t)nu * 32, 256), 256, 0, st>>>(nu, user0, (int)a->n_items, a->k, flag,
                                                                     cnt, citem, cscore, a->seen_ptr, a->seen_items,
                                                                     a->topk_items, a->topk_scores, fb, sc);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_EXHAUSTIVE, st);
      k_exhaustive<<<L.n_exh_ctas, 1024, ex_smem, st>>>(fb, sc, user0, (int)a->n_items, a->d, a->k, a->user_emb,
                                                       a->ld_user, a->user_ids, a->item_emb, a->ld_item, a->seen_ptr,
                                                       a->seen_items, scratch, a->topk_items, a->topk_scores);
      LGC_LAUNCH_CHECK();
    }
    if (a->stats) {
      k_add_stats<<<1, 1, 0, st>>>(sc, a->stats);
      LGC_LAUNCH_CHECK();
      LGC_CUDA(cudaMemsetAsync(&sc->n_groups, 0, 2 * sizeof(unsigned long long), st));
    }
  }
  return LGC_OK;
}
