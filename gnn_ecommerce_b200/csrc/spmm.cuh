// Internal interface of the CSR SpMM (shared by spmm.cu and train_step.cu).
#pragma once
#include <math.h>

#include "common.cuh"

namespace lgc {

enum EpiMode : int {
  EPI_PLAIN = 0,     // y = scale * (A x)[r] + beta * addend[r]          (LGConv; backward Horner)
  EPI_FWD_INIT = 1,  // acc = a0 * x[r] + a1 * (A x)[r];  y = (A x)[r] if y != null
  EPI_FWD_RMW = 2,   // acc += a1 * (A x)[r];             y = (A x)[r] if y != null
  EPI_ADAM = 3,      // g = scale * (A x)[r] + addend[r]; Adam update of p, m, v at row r
  EPI_FWD_FINAL = 4  // acc = (((hist0*ah0 + hist1*ah1) + ...) + (A x)[r]*a1): the whole layer mean of
                     // get_embedding in ONE pass over the stored layer tables, same rounding
                     // sequence as the reference's running sum (src/lightgcn.py:93,97)
};
constexpr int kMaxHist = 6;   // E0 + up to 5 stored layers (num_layers <= 6)

struct AdamScalars {
  float one_minus_beta1, beta2, one_minus_beta2, bc2_sqrt, eps, neg_step_size;
};

struct EpiArgs {
  float* y = nullptr;
  float* acc = nullptr;
  const float* xrow = nullptr;
  const float* addend = nullptr;
  // optional: bit r set <=> row r of `addend` may be non-zero (the sparse gradient tables of the
  // training step: <= 3 * batch rows). The rows kernel skips the addend stream of the other rows.
  const uint32_t* addend_mask = nullptr;
  // optional: bit s set <=> row s of the GATHERED table x may be non-zero (first backward layer: x is
  // the sparse gradient table itself). The sweep and the rows kernel skip the gathers of the other rows.
  const uint32_t* x_mask = nullptr;
  float a0 = 0.f, a1 = 0.f, scale = 1.f, beta = 0.f;
  float* p = nullptr;
  float* m = nullptr;
  float* v = nullptr;
  AdamScalars adam = {};
  const AdamScalars* adam_dev = nullptr;   // if set: scalars live in device memory (CUDA-graph replay)
  const float* hist[kMaxHist] = {};   // EPI_FWD_FINAL: x_0 (= E0), x_1, ..., x_{K-1}
  float ah[kMaxHist] = {};            //                their layer weights alpha_0 .. alpha_{K-1}
  int n_hist = 0;
};

// torch.optim.Adam (single-tensor path, no amsgrad / weight decay) on one element, with the
// operation order of ATen's CPU kernels (lerp_, mul_+addcmul_, sqrt/div/add_, addcdiv_).
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g,
                                            const AdamScalars& s) {
  m = fmaf(s.one_minus_beta1, __fsub_rn(g, m), m);
  v = __fadd_rn(__fmul_rn(v, s.beta2), __fmul_rn(__fmul_rn(s.one_minus_beta2, g), g));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(s.neg_step_size, m), denom));
}

// Same update for the SpMM epilogues, where the kernel is bound by issued instructions: the two
// IEEE divisions and the IEEE square root (~30 instructions with their slow paths) become
// sqrt.approx / rcp.approx + FMA (~8). Deviation from adam_update: a few ulp of the update term
// (<= 1e-6 relative to the step, ~1e-8 relative to the weight), inside the 1e-5 parity tolerance.
// `inv_bc2_sqrt` = 1 / s.bc2_sqrt (computed once per thread).
__device__ __forceinline__ void adam_update_fast(float& p, float& m, float& v, float g, const AdamScalars& s,
                                                 float inv_bc2_sqrt) {
  m = fmaf(s.one_minus_beta1, __fsub_rn(g, m), m);
  v = __fadd_rn(__fmul_rn(v, s.beta2), __fmul_rn(__fmul_rn(s.one_minus_beta2, g), g));
  float sq;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
  const float denom = fmaf(sq, inv_bc2_sqrt, s.eps);
  float rd;
  asm("rcp.approx.f32 %0, %1;" : "=f"(rd) : "f"(denom));
  p = fmaf(__fmul_rn(s.neg_step_size, m), rd, p);
}

static inline AdamScalars make_adam_scalars(double lr, double beta1, double beta2, double eps,
                                            int64_t step) {
  double bc1 = 1.0 - pow(beta1, (double)step);
  double bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars s;
  s.one_minus_beta1 = (float)(1.0 - beta1);
  s.beta2 = (float)beta2;
  s.one_minus_beta2 = (float)(1.0 - beta2);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.eps = (float)eps;
  s.neg_step_size = (float)(-(lr / bc1));
  return s;
}

// floats of scratch the heavy-row two-phase reduce needs for this graph and row width
size_t spmm_partials_floats(const lgc_graph* g, int ld);

// y/acc/... = epilogue(A_hat x). `partials` must hold spmm_partials_floats() floats.
int launch_spmm(const lgc_graph* g, int ld, const float* x, EpiMode mode, const EpiArgs& args,
                float* partials, cudaStream_t stream);

// ---- sweep path of the high-degree rows (sweep.cu). `sweep_get` builds the schedule for this row
// width on first use (synchronous, allocates inside the handle; never during stream capture: the
// workspace-size queries call it first) and returns nullptr when the rows stay with the chunked
// heavy-row kernel (unsupported width, node ids too large for the packed records, LGC_SWEEP=0).
const SweepSched* sweep_get(const lgc_graph* g, int ld);
bool sweep_has_rows(const SweepSched* s);          // false: no row qualified, nothing to launch
// The rows the sweep does not take, for the rows kernel (rows.cu): blocks of kRowsBlock consecutive
// rows; perm[b * kRowsBlock + j] = local index of the j-th of the block's blk_cnt[b] rows in
// descending-degree order; rec = (source, weight bits) per CSR entry. LGC_ROWS=0 keeps the round-1
// light-row kernel instead (no plan).
constexpr int kRowsBlock = 256;
struct RowPlan {
  int64_t n_blocks = 0, num_rows = 0, n_active = 0;   // n_active: rows the plan holds (0: nothing to launch)
  uint8_t* perm = nullptr;
  int32_t* blk_cnt = nullptr;
  int2* rec = nullptr;
};
const RowPlan* sweep_row_plan(const SweepSched* s);
void sweep_info(const SweepSched* s, lgc_plan_info* info);
bool rows_kernel_enabled();
int launch_rows(const lgc_graph* g, const RowPlan* plan, int ld, const float* x, EpiMode mode, const EpiArgs& args,
                cudaStream_t st);
size_t sweep_partial_slots(const SweepSched* s);
const int4* sweep_split_rows(const SweepSched* s, int64_t* n);
int launch_sweep(const lgc_graph* g, const SweepSched* s, int ld, const float* x, EpiMode mode,
                 const EpiArgs& args, float* partials, cudaStream_t st);

// out = sum_l alpha_l A^l x0 with K-1 scratch tables `xs` (see spmm.cu)
int propagate_chain(const lgc_graph* g, int ld, int K, const float* alpha, const float* x0, float* out,
                    float* const* xs, float* partials, cudaStream_t st);

}  // namespace lgc
