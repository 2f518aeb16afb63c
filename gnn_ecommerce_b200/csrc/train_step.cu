// Fused training step: one iteration of `TrainLightGCN.mini_batch_loop`
// (src/train_lightgcn.py:129-151) for pre-sampled triples, entirely on the device:
//   forward   out = sum_l alpha_l A^l E0              (K SpMM launches, running mean fused)
//   loss      BPR + layer-0 L2, sparse d(out)         (k_bpr)
//   backward  gE0 = alpha_0 G + A(alpha_1 G + A(...)) (Horner on the symmetric operator, K SpMMs)
//   update    dense Adam fused into the epilogue of the last backward SpMM
// G (= dL/d out) and Z (= alpha_0 G + L2 gradient) are dense [N, ld] buffers that are zero
// outside the <= 3*batch touched rows; they are re-zeroed sparsely at the end of the step.
#include "spmm.cuh"

namespace lgc {

int bpr_launch(int64_t num_nodes, int ld, int64_t batch, const int64_t* users, const int64_t* pos,
               const int64_t* neg, const float* out, const float* e0, double decay, float alpha0,
               float* grad_out, float* grad_e0, int32_t* touched, float* loss3, float* per_triple,
               float* rows, uint32_t* row_mask, int* bad, cudaStream_t st);

namespace {

struct StepLayout {
  size_t t;              // floats per table
  int n_x;               // scratch tables: the forward's stored layers, reused by the backward
  float* xs[kMaxHist];
  float *out, *g, *z, *partials, *per_triple, *rows;
  uint32_t* mask;        // bit per table row: set for the rows of G / Z that are not zero
  int32_t* touched;
  int* bad;
  size_t bytes;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

StepLayout layout(const lgc_graph* g, int ld, int num_layers, int64_t batch, void* base) {
  StepLayout L;
  L.t = (size_t)g->num_nodes * ld;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes, 256); return q; };
  L.g = (float*)take(L.t * 4);          // G and Z first: lgc_train_workspace_init zeroes them
  L.z = (float*)take(L.t * 4);
  L.mask = (uint32_t*)take(((size_t)g->num_nodes + 31) / 32 * 4 + 64);   // right behind G and Z: zeroed with them
  L.n_x = num_layers > 1 ? num_layers - 1 : 0;
  for (int i = 0; i < kMaxHist; ++i) L.xs[i] = i < L.n_x ? (float*)take(L.t * 4) : nullptr;
  L.out = (float*)take(L.t * 4);
  L.partials = (float*)take(spmm_partials_floats(g, ld) * 4);
  L.per_triple = (float*)take((size_t)batch * 2 * 4);
  L.touched = (int32_t*)take((size_t)batch * 3 * 4);
  L.rows = (float*)take((size_t)batch * 3 * ld * 4);
  L.bad = (int*)take(16);
  L.bytes = off;
  return L;
}

__global__ void k_zero_rows(const int32_t* __restrict__ rows, int64_t n_rows, int ld,
                            float* __restrict__ a, float* __restrict__ b, uint32_t* __restrict__ mask) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_rows || rows[i] < 0) return;
  if (lane == 0) mask[rows[i] >> 5] = 0u;             // every set bit of the word belongs to a touched row
  const size_t base = (size_t)rows[i] * ld;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = lane; c < ld / 4; c += 32) {
    st_f4(a + base + 4 * c, z);
    st_f4(b + base + 4 * c, z);
  }
}

__global__ void k_scale4(const float4* __restrict__ x, float4* __restrict__ y, float a, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    y[i] = make_float4(v.x * a, v.y * a, v.z * a, v.w * a);
  }
}

}  // namespace
}  // namespace lgc

using namespace lgc;

extern "C" size_t lgc_train_step_workspace_bytes(const lgc_graph_t* g, int ld, int num_layers,
                                                 int64_t batch) {
  return g ? layout(g, ld, num_layers, batch, nullptr).bytes : 0;
}

extern "C" int lgc_train_workspace_init(const lgc_graph_t* g, int ld, int num_layers, int64_t batch,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  LGC_REQUIRE(g && workspace, "null argument");
  if (workspace_bytes < lgc_train_step_workspace_bytes(g, ld, num_layers, batch)) {
    set_error("lgc_train_workspace_init: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  StepLayout L = layout(g, ld, num_layers, batch, workspace);
  LGC_CUDA(cudaMemsetAsync(L.g, 0, (size_t)((char*)L.mask - (char*)L.g) + ((size_t)g->num_nodes + 31) / 32 * 4 + 64,
                           (cudaStream_t)stream));
  return LGC_OK;
}

extern "C" int lgc_train_step(const lgc_graph_t* g, const lgc_train_step_args* a, void* stream) {
  LGC_REQUIRE(g && a, "null argument");
  LGC_REQUIRE(a->h_alpha && a->users && a->pos && a->neg && a->e0 && a->m && a->v && a->loss3 &&
                  a->workspace, "null field in lgc_train_step_args");
  LGC_REQUIRE(a->num_layers >= 0 && a->num_layers <= kMaxHist && a->batch > 0 && a->step >= 1,
              "bad num_layers/batch/step");
  LGC_REQUIRE(g->is_symmetric, "the fused step needs a symmetric normalised adjacency");
  LGC_REQUIRE(lgc_ld_supported(a->ld) && a->ld <= 256, "unsupported ld");
  if (a->workspace_bytes < lgc_train_step_workspace_bytes(g, a->ld, a->num_layers, a->batch)) {
    set_error("lgc_train_step: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  NvtxRange nvtx("lgc_train_step");
  cudaStream_t st = (cudaStream_t)stream;
  const int ld = a->ld, K = a->num_layers;
  const float* alpha = a->h_alpha;
  StepLayout L = layout(g, ld, K, a->batch, a->workspace);
  float* tmp[2] = {L.xs[0], L.xs[1]};       // backward ping-pong (K >= 3 has >= 2 scratch tables)
  int rc;

  // ---- forward: out = sum_l alpha_l A^l E0 (src/lightgcn.py:91-99)
  if (K == 0) {
    k_scale4<<<kNumSMs * 8, 256, 0, st>>>((const float4*)a->e0, (float4*)L.out, alpha[0], (int64_t)(L.t / 4));
    LGC_LAUNCH_CHECK();
  } else {
    rc = propagate_chain(g, ld, K, alpha, a->e0, L.out, L.xs, L.partials, st);
    if (rc) return rc;
  }

  // ---- loss + sparse gradients (src/lightgcn.py:123-125,279-286; src/utils_v2.py:193-211)
  LGC_CUDA(cudaMemsetAsync(L.bad, 0, sizeof(int), st));
  rc = bpr_launch(g->num_nodes, ld, a->batch, a->users, a->pos, a->neg, L.out, a->e0, a->decay, alpha[0],
                  L.g, L.z, L.touched, a->loss3, L.per_triple, L.rows, L.mask, L.bad, st);
  if (rc) return rc;

  // ---- backward (Horner) + Adam (src/train_lightgcn.py:146-147)
  AdamScalars as = make_adam_scalars(a->lr, a->beta1, a->beta2, a->eps, a->step);
  if (K == 0) {
    rc = lgc_adam_step((int64_t)L.t, a->e0, L.z, a->m, a->v, a->lr, a->beta1, a->beta2, a->eps, a->step, st);
    if (rc) return rc;
  } else {
    const float* cur = L.g;
    float scale = alpha[K];
    int flip = 0;
    for (int l = K - 1; l >= 1; --l) {          // h_l = alpha_l G + A h_{l+1}
      EpiArgs e;
      e.y = tmp[flip];
      e.scale = scale;
      e.beta = alpha[l];
      e.addend = L.g;
      e.addend_mask = L.mask;
      if (cur == L.g) e.x_mask = L.mask;        // first backward layer gathers from G itself
      rc = launch_spmm(g, ld, cur, EPI_PLAIN, e, L.partials, st);
      if (rc) return rc;
      cur = tmp[flip];
      flip ^= 1;
      scale = 1.f;
    }
    EpiArgs e;                                   // gE0 = Z + A h_1, consumed by Adam in place
    e.scale = scale;
    e.addend = L.z;
    e.addend_mask = L.mask;
    if (cur == L.g) e.x_mask = L.mask;
    e.p = a->e0; e.m = a->m; e.v = a->v;
    e.adam = as;
    rc = launch_spmm(g, ld, cur, EPI_ADAM, e, L.partials, st);
    if (rc) return rc;
  }
  k_zero_rows<<<(int)ceil_div(a->batch * 3 * 32, 256), 256, 0, st>>>(L.touched, a->batch * 3, ld, L.g, L.z, L.mask);
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}
