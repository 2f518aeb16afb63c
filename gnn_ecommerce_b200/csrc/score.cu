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

#include "common.cuh"

namespace lgc {
namespace {

constexpr int kUserBlock = 256;   // users per CTA step (two M=128 halves)
constexpr int kTileN = 128;       // items per MMA tile (N)
constexpr int kGroup = 16;        // items per stored maximum
constexpr int kCap = 64;          // exact candidates kept per user before falling back
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
// 32 lanes x 16 consecutive fp32 columns: thread i gets row (lane base + i)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// wait for this thread's outstanding tcgen05.ld; the registers are threaded through so the
// compiler cannot read them before the wait
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15])
               :
               : "memory");
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
// gmax16 : [n_tiles*4][u_pad] uint32 = two fp16 group maxima (groups 2p, 2p+1 of the tile)
// gmax128: [n_tiles][u_pad] fp16 tile maximum
__global__ void __launch_bounds__(kGemmThreads, 1)
k_score_gemm(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             int n_blocks, int n_tiles, int n_items, int katoms, int n_stages, int u_pad, float out_scale,
             uint32_t* __restrict__ gmax16, __half* __restrict__ gmax128) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t tile_bytes = 16384u * katoms;                 // one 128-row operand tile
  uint8_t* smem_a = smem;                                      // 2 tiles
  uint8_t* smem_b = smem + 2 * tile_bytes;                     // n_stages tiles
  uint64_t* bars = (uint64_t*)(smem_b + (size_t)n_stages * tile_bytes);
  // barrier slots: 0 a_full, 1 a_empty, 2..3 t_full, 4..5 t_empty, 6.. b_full[s], b_empty[s]
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 + 2 * 8);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar(0), 1); mbar_init(bar(1), 1);
    mbar_init(bar(2), 1); mbar_init(bar(3), 1);
    mbar_init(bar(4), 8); mbar_init(bar(5), 8);
    for (int s = 0; s < n_stages; ++s) { mbar_init(bar(6 + s), 1); mbar_init(bar(6 + 8 + s), 1); }
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
        mbar_wait(bar(1), pa ^ 1);
        mbar_expect_tx(bar(0), 2 * tile_bytes);
        for (int h = 0; h < 2; ++h)
          for (int ka = 0; ka < katoms; ++ka)
            tma_load_2d(smem_u32(smem_a + h * tile_bytes + ka * 16384), &map_a, ka * 64,
                        blk * kUserBlock + h * 128, bar(0));
        pa ^= 1;
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait(bar(6 + 8 + s), ph ^ 1);
          mbar_expect_tx(bar(6 + s), tile_bytes);
          for (int ka = 0; ka < katoms; ++ka)
            tma_load_2d(smem_u32(smem_b + (size_t)s * tile_bytes + ka * 16384), &map_b, ka * 64, t * kTileN,
                        bar(6 + s));
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread)
    if (lane == 0) {
      int s = 0, as = 0; uint32_t ph = 0, pa = 0, pt = 0;
      const int ksteps = katoms * 4;
      for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        mbar_wait(bar(0), pa);
        pa ^= 1;
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait(bar(6 + s), ph);
          mbar_wait(bar(4 + as), pt ^ 1);
          tc_fence_after();
          const uint32_t b_base = smem_u32(smem_b + (size_t)s * tile_bytes);
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const uint32_t a_base = smem_u32(smem_a + h * tile_bytes);
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256 + h * 128);
            for (int j = 0; j < ksteps; ++j) {
              const uint32_t koff = (uint32_t)(j >> 2) * 16384u + (uint32_t)(j & 3) * 32u;
              tc_mma_f16(d_tmem, umma_desc(a_base + koff), umma_desc(b_base + koff), kIdesc, j > 0);
            }
          }
          tc_commit(bar(6 + 8 + s));       // B stage free once these MMAs have read it
          tc_commit(bar(2 + as));          // accumulators ready for the epilogue
          if (++s == n_stages) { s = 0; ph ^= 1; }
          if (++as == 2) { as = 0; pt ^= 1; }
        }
        tc_commit(bar(1));                 // A tile free for the next user block
      }
    }
  } else {
    // ===== epilogue: warp w may touch TMEM lanes 32*(w%4)..+31
    const int q = warp & 3, h = (warp - 2) >> 2;
    int as = 0; uint32_t pt = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      const int user = blk * kUserBlock + h * 128 + q * 32 + lane;   // < u_pad by construction
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(bar(2 + as), pt);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + h * 128);
        const int item0 = t * kTileN;
        const bool ragged = item0 + kTileN > n_items;
        float gm[8];
        uint32_t va[16], vb[16];
        tc_ld16(taddr, va);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t (&cur)[16] = (c & 1) ? vb : va;
          uint32_t (&nxt)[16] = (c & 1) ? va : vb;
          tc_wait_ld(cur);
          if (c < 7) tc_ld16(taddr + 16 * (c + 1), nxt);
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(cur[i]);
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (item0 + c * 16 + i >= n_items) f[i] = -INFINITY;
          }
          const float m0 = max3(f[0], f[1], f[2]), m1 = max3(f[3], f[4], f[5]), m2 = max3(f[6], f[7], f[8]);
          const float m3 = max3(f[9], f[10], f[11]), m4 = max3(f[12], f[13], f[14]);
          gm[c] = fmaxf(max3(m0, m1, f[15]), max3(m2, m3, m4));
        }
        // all TMEM reads of this stage are complete: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(4 + as));
        if (++as == 2) { as = 0; pt ^= 1; }

        const float tm = fmaxf(max3(gm[0], gm[1], gm[2]), max3(max3(gm[3], gm[4], gm[5]), gm[6], gm[7]));
        uint32_t* g16 = gmax16 + (size_t)(t * 4) * u_pad + user;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const __half2 hh = __floats2half2_rn(gm[2 * p] * out_scale, gm[2 * p + 1] * out_scale);
          __stcs(g16 + (size_t)p * u_pad, *reinterpret_cast<const uint32_t*>(&hh));
        }
        gmax128[(size_t)t * u_pad + user] = __float2half_rn(tm * out_scale);
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

// ------------------------------------------------------------------------------------ threshold
__device__ __forceinline__ unsigned key_of(unsigned short bits) {   // monotone fp16 -> uint16
  return (bits & 0x8000u) ? (unsigned)(~bits & 0xFFFFu) : (unsigned)(bits | 0x8000u);
}
__device__ __forceinline__ unsigned short bits_of(unsigned key) {
  return (unsigned short)((key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu));
}

// CTA = 32 users; their tile maxima are staged in shared memory [n_tiles][32]; each warp then
// selects, for 4 users, the r-th largest (r = k + n_seen) by bisection on the 16 key bits.
__global__ void __launch_bounds__(256)
k_threshold(const __half* __restrict__ gmax128, const uint32_t* __restrict__ gmax16, int use_groups, int n_tiles,
            int u_pad, int n_users, int64_t user0, int k,
            const int64_t* __restrict__ seen_ptr, const float* __restrict__ anorm,
            ScoreScalars* sc, int kp, float out_scale, float* __restrict__ thr_grp,
            float* __restrict__ thr_exact, uint8_t* __restrict__ flag, int32_t* __restrict__ fb_users) {
  extern __shared__ unsigned short s_keys[];   // [n_sel][32]
  const int u0 = blockIdx.x * 32;
  // few tiles (small item sets): select among the 16-item group maxima instead -- the same bound
  const int n_sel = use_groups ? n_tiles * 8 : n_tiles;
  if (use_groups) {
    for (int i = threadIdx.x; i < n_sel * 32; i += 256) {
      const int t = i >> 5, j = i & 31;
      const uint32_t w = gmax16[(size_t)(t >> 1) * u_pad + u0 + j];
      s_keys[i] = (unsigned short)key_of((unsigned short)((t & 1) ? (w >> 16) : (w & 0xFFFFu)));
    }
  } else {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(gmax128);
    for (int i = threadIdx.x; i < n_sel * 32; i += 256) {
      const int t = i >> 5, j = i & 31;
      s_keys[i] = (unsigned short)key_of(src[(size_t)t * u_pad + u0 + j]);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ea = scale_exp(sc->amax_bits), eb = scale_exp(sc->bmax_bits);
  const float bnorm = __uint_as_float(sc->bnorm_bits);
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
    unsigned prefix = 0;
    for (int bit = 15; bit >= 0; --bit) {
      const unsigned cand = prefix | (1u << bit);
      int cnt = 0;
      for (int t = lane; t < n_sel; t += 32) cnt += (s_keys[t * 32 + j] >= cand) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt >= r) prefix = cand;
    }
    if (lane == 0) {
      const float T = __half2float(__ushort_as_half(bits_of(prefix)));      // stored units
      const float eps = (kEpsRel * anorm[u] * bnorm + (float)kp * 0.001953125f) * out_scale;
      flag[u] = 0;
      thr_grp[u] = T - 2.f * eps;
      // back to true units: stored = score * 2^(ea+eb) * out_scale
      thr_exact[u] = scalbnf((T - eps) / out_scale, -(ea + eb));
    }
  }
}

// ------------------------------------------------------------------------------------ exact re-scoring
// Fixed summation order shared by the fast path and the exhaustive path (identical scores):
// half `hf` takes the float4 chunks hf, hf+2, ...; the total is half0 + half1.
__device__ __forceinline__ float dot_half(const float4* __restrict__ a, const float4* __restrict__ b, int d4, int hf) {
  float s = 0.f;
  for (int c = hf; c < d4; c += 2) {
    const float4 x = a[c], y = b[c];
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}

constexpr int kRescoreBatch = 1024;   // users scanned per CTA iteration

// grid (n_tiles, user splits). The CTA keeps its 128 item rows (fp32) in shared memory, scans the
// stored group maxima of its users against their thresholds, and re-scores the hits exactly.
__global__ void __launch_bounds__(256)
k_rescore(const uint32_t* __restrict__ gmax16, int u_pad, int n_users, int64_t user0, int n_items, int d,
          const float* __restrict__ user_emb, int ld_user, const int64_t* __restrict__ user_ids,
          const float* __restrict__ item_emb, int ld_item, const float* __restrict__ thr_grp,
          const float* __restrict__ thr_exact, const int64_t* __restrict__ seen_ptr,
          const int64_t* __restrict__ seen_items, int* __restrict__ cand_cnt, int32_t* __restrict__ cand_item,
          float* __restrict__ cand_score, ScoreScalars* __restrict__ sc) {
  extern __shared__ __align__(16) uint8_t smem_rs[];
  const int d4 = (d + 3) >> 2;
  const int row_f4 = d4 + 1;                          // +16 B per row: conflict-free float4 reads
  float4* s_items = reinterpret_cast<float4*>(smem_rs);                 // [128][row_f4]
  float4* s_user = s_items + 128 * row_f4;                              // [8 warps][d4]
  int* s_queue = reinterpret_cast<int*>(s_user + 8 * d4);               // [kRescoreBatch * 8]
  __shared__ int s_qn;

  const int tile = blockIdx.x, item0 = tile * kTileN;
  for (int i = threadIdx.x; i < 128 * d4; i += 256) {
    const int r = i / d4, c = i % d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (item0 + r < n_items) {
      const float* p = item_emb + (size_t)(item0 + r) * ld_item + 4 * c;
      if (4 * c + 3 < d) v = *reinterpret_cast<const float4*>(p);
      else { v.x = p[0]; if (4 * c + 1 < d) v.y = p[1]; if (4 * c + 2 < d) v.z = p[2]; }
    }
    s_items[r * row_f4 + c] = v;
  }
  const int per_split = round_up((u_pad + gridDim.y - 1) / gridDim.y, kRescoreBatch);
  const int u_beg = blockIdx.y * per_split, u_end = min(n_users, u_beg + per_split);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long my_groups = 0, my_emit = 0;

  for (int base = u_beg; base < u_end; base += kRescoreBatch) {
    if (threadIdx.x == 0) s_qn = 0;
    __syncthreads();
    // ---- scan
#pragma unroll
    for (int j = 0; j < kRescoreBatch / 256; ++j) {
      const int u = base + threadIdx.x + 256 * j;
      if (u < u_end) {
        const float thr = thr_grp[u];
        uint32_t w[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) w[p] = __ldcs(gmax16 + (size_t)(tile * 4 + p) * u_pad + u);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[p]));
          if (f.x >= thr) s_queue[atomicAdd(&s_qn, 1)] = (u << 3) | (2 * p);
          if (f.y >= thr) s_queue[atomicAdd(&s_qn, 1)] = (u << 3) | (2 * p + 1);
        }
      }
    }
    __syncthreads();
    const int qn = s_qn;
    // ---- exact scores of the hit groups: one warp per (user, group)
    for (int e = warp; e < qn; e += 8) {
      const int ent = s_queue[e], u = ent >> 3, g = ent & 7;
      const int64_t uid = user_ids ? user_ids[user0 + u] : user0 + u;
      const float* urow = user_emb + (size_t)uid * ld_user;
      float4* su = s_user + warp * d4;
      __syncwarp();
      for (int c = lane; c < d4; c += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * c + 3 < d) v = *reinterpret_cast<const float4*>(urow + 4 * c);
        else { v.x = urow[4 * c]; if (4 * c + 1 < d) v.y = urow[4 * c + 1]; if (4 * c + 2 < d) v.z = urow[4 * c + 2]; }
        su[c] = v;
      }
      __syncwarp();
      const int it = lane & 15, hf = lane >> 4;
      const int local = g * kGroup + it, item = item0 + local;
      float s = dot_half(su, s_items + local * row_f4, d4, hf);
      const float other = __shfl_xor_sync(0xffffffffu, s, 16);
      s = hf == 0 ? s + other : other + s;           // half0 + half1 on every lane
      if (hf == 0 && item < n_items && s >= thr_exact[u]) {
        bool seen = false;
        if (seen_ptr) {
          for (int64_t q = seen_ptr[user0 + u]; q < seen_ptr[user0 + u + 1]; ++q) seen |= (seen_items[q] == item);
        }
        if (!seen) {
          const int slot = atomicAdd(&cand_cnt[u], 1);
          if (slot < kCap) {
            cand_item[(size_t)u * kCap + slot] = item;
            cand_score[(size_t)u * kCap + slot] = s;
          }
          ++my_emit;
        }
      }
    }
    if (threadIdx.x == 0) my_groups += qn;
  }
  if (my_emit) atomicAdd(&sc->n_emitted, my_emit);
  if (threadIdx.x == 0 && my_groups) atomicAdd(&sc->n_groups, my_groups);
}

// ------------------------------------------------------------------------------------ final selection
// Strict total order: a before b iff (a.s > b.s) or (a.s == b.s and a.item < b.item).
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// One warp per user. Entries = exact candidates + the user's seen items with masked score 0.0
// (reference: pred * (1 - mask), src/lightgcn.py:175). Round r picks the first entry strictly
// after the previous pick, so duplicates collapse and nothing is mutated.
__global__ void __launch_bounds__(256)
k_select(int n_users, int64_t user0, int n_items, int k, uint8_t* flag, const int* __restrict__ cand_cnt, const int32_t* __restrict__ cand_item,
         const float* __restrict__ cand_score, const int64_t* __restrict__ seen_ptr,
         const int64_t* __restrict__ seen_items, int64_t* __restrict__ topk_items,
         float* __restrict__ topk_scores, int32_t* __restrict__ fb_users, ScoreScalars* __restrict__ sc) {
  const int u = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (u >= n_users || flag[u]) return;
  const int cnt = cand_cnt[u];
  if (cnt > kCap) {
    if (lane == 0) { flag[u] = 1; fb_users[atomicAdd(&sc->fb_count, 1)] = u; }
    return;
  }
  const int64_t sb = seen_ptr ? seen_ptr[user0 + u] : 0, se = seen_ptr ? seen_ptr[user0 + u + 1] : 0;
  const int n_seen = (int)(se - sb), total = cnt + n_seen;
  float ps = INFINITY; int pi = -1;
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY; int bi = 0x7fffffff;
    for (int e = lane; e < total; e += 32) {
      float s; int it;
      if (e < cnt) { s = cand_score[(size_t)u * kCap + e]; it = cand_item[(size_t)u * kCap + e]; }
      else {
        const int64_t si = seen_items[sb + (e - cnt)];
        if (si < 0 || si >= n_items) continue;
        s = 0.f; it = (int)si;
      }
      if (before(ps, pi, s, it) && before(s, it, bs, bi)) { bs = s; bi = it; }
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
__global__ void __launch_bounds__(256)
k_exhaustive(const int32_t* __restrict__ fb_users, const ScoreScalars* __restrict__ sc, int64_t user0,
             int n_items, int d, int k, const float* __restrict__ user_emb, int ld_user,
             const int64_t* __restrict__ user_ids, const float* __restrict__ item_emb, int ld_item,
             const int64_t* __restrict__ seen_ptr, const int64_t* __restrict__ seen_items,
             float* __restrict__ scratch, int64_t* __restrict__ topk_items, float* __restrict__ topk_scores) {
  extern __shared__ __align__(16) uint8_t smem_ex[];
  const int d4 = (d + 3) >> 2;
  float4* s_user = reinterpret_cast<float4*>(smem_ex);
  __shared__ float s_bs[8];
  __shared__ int s_bi[8];
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
    for (int c = threadIdx.x; c < d4; c += 256) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = urow[4 * c];
      if (4 * c + 1 < d) v.y = urow[4 * c + 1];
      if (4 * c + 2 < d) v.z = urow[4 * c + 2];
      if (4 * c + 3 < d) v.w = urow[4 * c + 3];
      s_user[c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_items; i += 256) {
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
      for (int64_t q = seen_ptr[user0 + u] + threadIdx.x; q < seen_ptr[user0 + u + 1]; q += 256) {
        const int64_t si = seen_items[q];
        if (si >= 0 && si < n_items) my[si] = 0.f;          // multiplicative mask
      }
    }
    if (threadIdx.x == 0) { s_ps = INFINITY; s_pi = -1; }
    __syncthreads();
    for (int r = 0; r < k; ++r) {
      const float ps = s_ps; const int pi = s_pi;
      float bs = -INFINITY; int bi = 0x7fffffff;
      for (int i = threadIdx.x; i < n_items; i += 256) {
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
        for (int w = 1; w < 8; ++w)
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
  int kp, katoms, n_stages, n_tiles, i_pad, chunk, chunk_pad;
  size_t gemm_smem;
  size_t off_scal, off_b16, off_a16, off_anorm, off_g16, off_g128, off_thr_grp, off_thr_exact, off_flag,
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
  int ns = (int)((budget - 2 * tile_bytes) / tile_bytes);
  L->n_stages = std::max(1, std::min(ns, 6));
  L->gemm_smem = 1024 + (2 + (size_t)L->n_stages) * tile_bytes + 256;
  L->i_pad = round_up(n_items, kTileN);
  L->n_tiles = L->i_pad / kTileN;
  // bytes per user of the chunk-sized buffers
  const size_t per_user = (size_t)L->kp * 2 + 4 + (size_t)L->n_tiles * 4 * 4 + (size_t)L->n_tiles * 2 + 4 + 4 + 1 +
                          4 + 4 + (size_t)kCap * 8;
  const size_t fixed = (size_t)L->i_pad * L->kp * 2 + (size_t)296 * n_items * 4 + (1 << 16);
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
  L->off_a16 = take(cp * L->kp * 2);
  L->off_anorm = take(cp * 4);
  L->off_g16 = take((size_t)L->n_tiles * 4 * cp * 4);
  L->off_g128 = take((size_t)L->n_tiles * cp * 2);
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
  CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim,
                                      gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    const char* s = nullptr;
    cuGetErrorString(r, &s);
    set_error(std::string("cuTensorMapEncodeTiled: ") + (s ? s : "?"));
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
  float* anorm = (float*)(ws + L.off_anorm);
  uint32_t* g16 = (uint32_t*)(ws + L.off_g16);
  __half* g128 = (__half*)(ws + L.off_g128);
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

  static bool attr_done = false;
  if (!attr_done) {
    LGC_CUDA(cudaFuncSetAttribute(k_score_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_threshold, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LGC_CUDA(cudaFuncSetAttribute(k_rescore, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }

  CUtensorMap map_a, map_b;
  int rc = encode_map(&map_a, a16, L.kp, L.chunk_pad);
  if (rc) return rc;
  rc = encode_map(&map_b, b16, L.kp, L.i_pad);
  if (rc) return rc;

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
        a->item_emb, a->ld_item, a->d, nullptr, 0, a->n_items, L.i_pad, L.kp, &sc->bmax_bits, b16, nullptr,
        &sc->bnorm_bits);
    LGC_LAUNCH_CHECK();
  }

  const int use_groups = L.n_tiles < 256 ? 1 : 0;
  const size_t thr_smem = (size_t)L.n_tiles * (use_groups ? 8 : 1) * 32 * 2;
  if (thr_smem > 200 * 1024) {
    set_error("lgc_score_topk: n_items above 409600 is not supported yet");
    return LGC_ERR_UNSUPPORTED;
  }
  const int d4 = (a->d + 3) / 4;
  const size_t rs_smem = (size_t)128 * (d4 + 1) * 16 + (size_t)8 * d4 * 16 + (size_t)kRescoreBatch * 8 * 4;
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
      k_score_gemm<<<grid, kGemmThreads, L.gemm_smem, st>>>(map_a, map_b, n_blocks, L.n_tiles, (int)a->n_items,
                                                           L.katoms, L.n_stages, L.chunk_pad, out_scale, g16, g128);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_SELECT, st);
      k_threshold<<<nu_pad / 32, 256, thr_smem, st>>>(g128, g16, use_groups, L.n_tiles, L.chunk_pad, nu, user0, a->k, a->seen_ptr,
                                                     anorm, sc, L.kp, out_scale, thr_grp, thr_exact, flag, fb);
      LGC_LAUNCH_CHECK();
    }
    {
      ProfScope ps(PROF_SCORE_RESCORE, st);
      int splits = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(kNumSMs * 8, L.n_tiles),
                                                               ceil_div(nu, kRescoreBatch)));
      dim3 grid(L.n_tiles, splits);
      k_rescore<<<grid, 256, rs_smem, st>>>(g16, L.chunk_pad, nu, user0, (int)a->n_items, a->d, a->user_emb,
                                            a->ld_user, a->user_ids, a->item_emb, a->ld_item, thr_grp, thr_exact,
                                            a->seen_ptr, a->seen_items, cnt, citem, cscore, sc);
      LGC_LAUNCH_CHECK();
      k_select<<<(int)ceil_div((int64_t)nu * 32, 256), 256, 0, st>>>(nu, user0, (int)a->n_items, a->k, flag,
                                                                     cnt, citem, cscore, a->seen_ptr, a->seen_items,
                                                                     a->topk_items, a->topk_scores, fb, sc);
      LGC_LAUNCH_CHECK();
      k_exhaustive<<<L.n_exh_ctas, 256, ex_smem, st>>>(fb, sc, user0, (int)a->n_items, a->d, a->k, a->user_emb,
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
