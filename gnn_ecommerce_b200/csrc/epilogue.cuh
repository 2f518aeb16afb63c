// Fused SpMM epilogues on W-float vectors (W = 1, 2 or 4), shared by the kernels whose lanes do not
// own whole float4 columns (sweep.cu: 16 lanes per row, ld / 16 floats per lane; rows.cu). Same
// arithmetic and rounding sequence as the float4 versions in spmm.cu (what the reference runs as
// separate ATen passes: the running layer mean `out = out + x * alpha`, src/lightgcn.py:93,97; the
// backward Horner add; torch.optim.Adam, src/train_lightgcn.py:147).
#pragma once
#include "spmm.cuh"

namespace lgc {

template <int W>
__device__ __forceinline__ void ldv(const float* p, float (&r)[W]) {
  if constexpr (W == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else if constexpr (W == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    r[0] = v.x; r[1] = v.y;
  } else {
    r[0] = *p;
  }
}
template <int W>
__device__ __forceinline__ void stv(float* p, const float (&r)[W]) {
  if constexpr (W == 4) *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  else if constexpr (W == 2) *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
  else *p = r[0];
}
// read-only (non-coherent) load: gathers of neighbour rows, cached in L1 / L2
template <int W>
__device__ __forceinline__ void ldv_nc(const float* p, float* r) {
  if constexpr (W == 4) {
    asm("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "l"(p));
  } else if constexpr (W == 2) {
    asm("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "l"(p));
  } else {
    asm("ld.global.nc.f32 %0, [%1];" : "=f"(r[0]) : "l"(p));
  }
}
// streaming (touched once per launch): evict-first, no L1 allocation
template <int W>
__device__ __forceinline__ void ldv_stream(const float* p, float (&r)[W]) {
  if constexpr (W == 4) {
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "l"(p));
  } else if constexpr (W == 2) {
    asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "l"(p));
  } else {
    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(r[0]) : "l"(p));
  }
}
template <int W>
__device__ __forceinline__ void stv_stream(float* p, const float (&r)[W]) {
  if constexpr (W == 4) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
  } else if constexpr (W == 2) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(r[0]), "f"(r[1]) : "memory");
  } else {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(r[0]) : "memory");
  }
}

// Epilogue operands that do not depend on the SpMM sum; loaded before the gathers so that their
// latency overlaps the gather latency.
template <int MODE, int W>
struct EpiPre {
  static constexpr int N = MODE == EPI_ADAM ? 4 : (MODE == EPI_FWD_FINAL ? kMaxHist : 1);
  float q[N][W];
};

// `addend_on` = false: the addend row is known to be zero (EpiArgs::addend_mask): not read
template <int MODE, int W>
__device__ __forceinline__ void epi_preload_w(const EpiArgs& a, size_t off, EpiPre<MODE, W>& p,
                                              bool addend_on = true) {
  if constexpr (MODE == EPI_PLAIN) {
#pragma unroll
    for (int i = 0; i < W; ++i) p.q[0][i] = 0.f;
    if (a.addend && addend_on) ldv_stream<W>(a.addend + off, p.q[0]);
  } else if constexpr (MODE == EPI_FWD_INIT) {
    ldv_stream<W>(a.xrow + off, p.q[0]);
  } else if constexpr (MODE == EPI_FWD_RMW) {
    ldv_stream<W>(a.acc + off, p.q[0]);
  } else if constexpr (MODE == EPI_ADAM) {
#pragma unroll
    for (int i = 0; i < W; ++i) p.q[0][i] = 0.f;
    if (addend_on) ldv_stream<W>(a.addend + off, p.q[0]);
    ldv_stream<W>(a.p + off, p.q[1]);
    ldv_stream<W>(a.m + off, p.q[2]);
    ldv_stream<W>(a.v + off, p.q[3]);
  } else {  // EPI_FWD_FINAL
#pragma unroll
    for (int i = 0; i < kMaxHist; ++i)
      if (i < a.n_hist) ldv_stream<W>(a.hist[i] + off, p.q[i]);
  }
}

template <int MODE, int W>
__device__ __forceinline__ void epi_finish_w(const EpiArgs& a, size_t off, const float (&s)[W],
                                             const EpiPre<MODE, W>& p) {
  float r[W];
  if constexpr (MODE == EPI_PLAIN) {
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = a.scale * s[i];
    if (a.addend) {
#pragma unroll
      for (int i = 0; i < W; ++i) r[i] = fmaf(a.beta, p.q[0][i], r[i]);
    }
    stv<W>(a.y + off, r);
  } else if constexpr (MODE == EPI_FWD_INIT) {
    if (a.y) stv<W>(a.y + off, s);
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = __fadd_rn(__fmul_rn(p.q[0][i], a.a0), __fmul_rn(s[i], a.a1));
    stv_stream<W>(a.acc + off, r);
  } else if constexpr (MODE == EPI_FWD_RMW) {
    if (a.y) stv<W>(a.y + off, s);
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = __fadd_rn(p.q[0][i], __fmul_rn(s[i], a.a1));
    stv_stream<W>(a.acc + off, r);
  } else if constexpr (MODE == EPI_ADAM) {
    const AdamScalars ad = a.adam_dev ? *a.adam_dev : a.adam;
    const float ib = __frcp_rn(ad.bc2_sqrt);
    float pp[W], mm[W], vv[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      pp[i] = p.q[1][i]; mm[i] = p.q[2][i]; vv[i] = p.q[3][i];
      adam_update_fast(pp[i], mm[i], vv[i], fmaf(a.scale, s[i], p.q[0][i]), ad, ib);
    }
    stv<W>(a.p + off, pp);
    stv_stream<W>(a.m + off, mm);
    stv_stream<W>(a.v + off, vv);
  } else {  // EPI_FWD_FINAL: (((x0*a0 + x1*a1) + ...) + s*a_K), the reference's running sum
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = __fmul_rn(p.q[0][i], a.ah[0]);
#pragma unroll
    for (int h = 1; h < kMaxHist; ++h)
      if (h < a.n_hist) {
#pragma unroll
        for (int i = 0; i < W; ++i) r[i] = __fadd_rn(r[i], __fmul_rn(p.q[h][i], a.ah[h]));
      }
#pragma unroll
    for (int i = 0; i < W; ++i) r[i] = __fadd_rn(r[i], __fmul_rn(s[i], a.a1));
    stv_stream<W>(a.acc + off, r);
  }
}

template <int MODE, int W>
__device__ __forceinline__ void epilogue_w(const EpiArgs& a, size_t off, const float (&s)[W]) {
  EpiPre<MODE, W> p;
  epi_preload_w<MODE, W>(a, off, p);
  epi_finish_w<MODE, W>(a, off, s, p);
}

}  // namespace lgc
