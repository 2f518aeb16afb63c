// Fused BPR step pieces: pair scores, loss + gradients, dense Adam.
// Replaces the ATen chain of src/lightgcn.py:123-125 (gather, mul, reduce), :279-286
// (logsigmoid mean), src/utils_v2.py:193-211 (layer-0 L2 term), their autograd backward
// (src/train_lightgcn.py:146) and torch.optim.Adam (src/train_lightgcn.py:58,147).
#include "spmm.cuh"

namespace lgc {
namespace {

constexpr int kMaxVecPerLane = 2;  // ld <= 256 floats

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

__global__ void k_pair_scores(const float* __restrict__ out, int ld, const int64_t* __restrict__ pairs,
                              int64_t n_pairs, float* __restrict__ score) {
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n_pairs) return;
  const float* a = out + (size_t)pairs[j] * ld;
  const float* b = out + (size_t)pairs[n_pairs + j] * ld;
  float s = 0.f;
  for (int c = lane; c < ld / 4; c += 32) s += dot4(ldg_f4(a + 4 * c), ldg_f4(b + 4 * c));
  s = warp_sum(s);
  if (lane == 0) score[j] = s;
}

// One warp per (user, pos, neg) triple: pair scores, loss terms, and the three gradient rows of
// d loss / d out written to `rows` ([3 * batch, ld]: row 3t = user, 3t+1 = pos, 3t+2 = neg) with their
// node ids in `touched`. No atomics: k_bpr_scatter adds the rows to the tables in a fixed order.
__global__ void k_bpr(int64_t num_nodes, int ld, int64_t batch, const int64_t* __restrict__ users,
                      const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                      const float* __restrict__ out, const float* __restrict__ e0, float inv_batch,
                      float* __restrict__ rows, int32_t* __restrict__ touched,
                      float* __restrict__ per_triple, int* __restrict__ bad) {
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= batch) return;
  const int64_t u = users[t], p = pos[t], n = neg[t];
  if (u < 0 || u >= num_nodes || p < 0 || p >= num_nodes || n < 0 || n >= num_nodes) {
    // the reference raises IndexError / a device assert: flag it (k_bpr_reduce turns the losses into NaN)
    if (lane == 0) { atomicExch(bad, 1); per_triple[2 * t] = 0.f; per_triple[2 * t + 1] = 0.f; }
    if (lane < 3) touched[3 * t + lane] = -1;
    return;
  }
  const int vec = ld / 4;
  float4 ou[kMaxVecPerLane], op[kMaxVecPerLane], on[kMaxVecPerLane];
  float sp = 0.f, sn = 0.f, sq = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < vec) {
      ou[k] = ldg_f4(out + (size_t)u * ld + 4 * c);
      op[k] = ldg_f4(out + (size_t)p * ld + 4 * c);
      on[k] = ldg_f4(out + (size_t)n * ld + 4 * c);
      sp += dot4(ou[k], op[k]);
      sn += dot4(ou[k], on[k]);
      float4 a = ldg_f4(e0 + (size_t)u * ld + 4 * c), b = ldg_f4(e0 + (size_t)p * ld + 4 * c),
             d = ldg_f4(e0 + (size_t)n * ld + 4 * c);
      sq += dot4(a, a) + dot4(b, b) + dot4(d, d);
    }
  }
  sp = warp_sum(sp); sn = warp_sum(sn); sq = warp_sum(sq);
  const float x = sp - sn;                                    // s+ - s-
  // -logsigmoid(x) = softplus(-x) = max(-x, 0) + log1p(exp(-|x|))
  const float loss = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
  const float sig = 1.f / (1.f + expf(x));                    // sigmoid(-x)
  const float coef = -sig * inv_batch;                        // d loss / d s+  (= - d loss / d s-)
  if (lane == 0) { per_triple[2 * t] = loss; per_triple[2 * t + 1] = sq; }
  if (lane < 3) touched[3 * t + lane] = (int32_t)(lane == 0 ? u : (lane == 1 ? p : n));
  float* r0 = rows + (size_t)(3 * t) * ld;
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < vec) {
      const float4 gu = make_float4(coef * (op[k].x - on[k].x), coef * (op[k].y - on[k].y),
                                    coef * (op[k].z - on[k].z), coef * (op[k].w - on[k].w));
      const float4 gp = make_float4(coef * ou[k].x, coef * ou[k].y, coef * ou[k].z, coef * ou[k].w);
      st_f4(r0 + 4 * c, gu);
      st_f4(r0 + ld + 4 * c, gp);
      st_f4(r0 + 2 * ld + 4 * c, make_float4(-gp.x, -gp.y, -gp.z, -gp.w));
    }
  }
}

// grad_out[row] += sum of the gradient rows of `row`; grad_e0[row] += sum of (alpha0 * gradient row +
// (decay / batch) * e0[row]) -- one term per occurrence, i.e. the layer-0 L2 term with multiplicity
// (src/utils_v2.py:193-211). Deterministic: one warp per gradient row j; it is the leader of its node
// if no earlier row carries the same id, and a leader adds all rows of that node in input order
// (users first, then positives, then negatives of a triple; triples in batch order). The reference's
// scatter in autograd's index_put backward is likewise sequential on CPU; the round-1 kernel used
// float4 atomics, whose order changed from run to run. Also marks the node in `row_mask` (bit per
// row of the tables: the SpMM epilogues skip the addend rows whose bit is clear).
__global__ void k_bpr_scatter(int n, int ld, const int32_t* __restrict__ touched, const float* __restrict__ rows,
                              const float* __restrict__ e0, float decay_over_batch, float alpha0,
                              float* __restrict__ grad_out, float* __restrict__ grad_e0,
                              uint32_t* __restrict__ row_mask) {
  const int j = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const int me = touched[j];
  if (me < 0) return;
  bool dup = false;
  for (int q = lane; q < j; q += 32) dup |= (touched[q] == me);
  if (__any_sync(0xffffffffu, dup)) return;
  const int vec = ld / 4;
  float4 g[kMaxVecPerLane], z[kMaxVecPerLane], e[kMaxVecPerLane];
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    g[k] = z[k] = e[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane + 32 * k;
    if (c < vec && grad_e0) e[k] = ldg_f4(e0 + (size_t)me * ld + 4 * c);
  }
  const float r = decay_over_batch;
  for (int base = j; base < n; base += 32) {
    const int q = base + lane;
    unsigned hit = __ballot_sync(0xffffffffu, q < n && touched[q] == me);
    while (hit) {
      const int rr = base + __ffs(hit) - 1;
      hit &= hit - 1;
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int c = lane + 32 * k;
        if (c < vec) {
          const float4 x = ldg_f4(rows + (size_t)rr * ld + 4 * c);
          g[k] = add4(g[k], x);
          z[k] = add4(z[k], make_float4(fmaf(alpha0, x.x, r * e[k].x), fmaf(alpha0, x.y, r * e[k].y),
                                        fmaf(alpha0, x.z, r * e[k].z), fmaf(alpha0, x.w, r * e[k].w)));
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < vec) {
      float* dg = grad_out + (size_t)me * ld + 4 * c;
      st_f4(dg, add4(ld_f4(dg), g[k]));
      if (grad_e0) {
        float* dz = grad_e0 + (size_t)me * ld + 4 * c;
        st_f4(dz, add4(ld_f4(dz), z[k]));
      }
    }
  }
  if (row_mask && lane == 0) atomicOr(&row_mask[me >> 5], 1u << (me & 31));
}

// Deterministic reduction of the per-triple terms (fixed tree, double accumulators).
__global__ void k_bpr_reduce(const float* __restrict__ per_triple, int64_t batch, double decay,
                             const int* __restrict__ bad, float* __restrict__ loss3) {
  __shared__ double s_loss[256], s_sq[256];
  double a = 0.0, b = 0.0;
  for (int64_t t = threadIdx.x; t < batch; t += 256) { a += per_triple[2 * t]; b += per_triple[2 * t + 1]; }
  s_loss[threadIdx.x] = a; s_sq[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) { s_loss[threadIdx.x] += s_loss[threadIdx.x + o]; s_sq[threadIdx.x] += s_sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float bpr = (float)(s_loss[0] / (double)batch);
    float reg = (float)(0.5 * s_sq[0] / (double)batch * decay);
    if (*bad) bpr = reg = __int_as_float(0x7fc00000);      // an id outside [0, num_nodes): NaN, not a silently biased loss
    loss3[0] = bpr; loss3[1] = reg; loss3[2] = bpr + reg;
  }
}

// table[idx[j]] += rows[j] for j < n with duplicates summed in input order (deterministic: the
// multi-GPU step keeps REPLICATED item tables that must stay bit-identical on every rank, which
// atomics cannot guarantee). One warp per input row j: it is the leader of its index if no
// earlier row carries the same index; a leader adds all rows of that index, in order.
// idx < 0 entries are skipped (rows that belong to another rank).
__global__ void k_scatter_add_rows(int n, int ld, const int64_t* __restrict__ idx, const float* __restrict__ rows,
                                   float* __restrict__ table) {
  const int j = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const int64_t me = idx[j];
  if (me < 0) return;
  bool dup = false;
  for (int q = lane; q < j; q += 32) dup |= (idx[q] == me);
  if (__any_sync(0xffffffffu, dup)) return;
  const int vec = ld / 4;
  float4 acc[kMaxVecPerLane];
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = j; base < n; base += 32) {
    const int q = base + lane;
    unsigned hit = __ballot_sync(0xffffffffu, q < n && idx[q] == me);
    while (hit) {
      const int r = base + __ffs(hit) - 1;
      hit &= hit - 1;
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int c = lane + 32 * k;
        if (c < vec) acc[k] = add4(acc[k], ldg_f4(rows + (size_t)r * ld + 4 * c));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < vec) {
      float* dst = table + (size_t)me * ld + 4 * c;
      st_f4(dst, add4(ld_f4(dst), acc[k]));
    }
  }
}

__global__ void k_adam(int64_t n4, float4* __restrict__ p, const float4* __restrict__ g,
                       float4* __restrict__ m, float4* __restrict__ v, AdamScalars s_host,
                       const AdamScalars* __restrict__ s_dev) {
  const AdamScalars s = s_dev ? *s_dev : s_host;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = __ldcs(g + i), mm = __ldcs(m + i), vv = __ldcs(v + i);
    adam_update(pp.x, mm.x, vv.x, gg.x, s);
    adam_update(pp.y, mm.y, vv.y, gg.y, s);
    adam_update(pp.z, mm.z, vv.z, gg.z, s);
    adam_update(pp.w, mm.w, vv.w, gg.w, s);
    p[i] = pp; __stcs(m + i, mm); __stcs(v + i, vv);
  }
}

__global__ void k_adam_tail(int64_t beg, int64_t n, float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, AdamScalars s_host,
                            const AdamScalars* __restrict__ s_dev) {
  const AdamScalars s = s_dev ? *s_dev : s_host;
  int64_t i = beg + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) adam_update(p[i], m[i], v[i], g[i], s);
}

}  // namespace

int bpr_launch(int64_t num_nodes, int ld, int64_t batch, const int64_t* users, const int64_t* pos,
               const int64_t* neg, const float* out, const float* e0, double decay, float alpha0,
               float* grad_out, float* grad_e0, int32_t* touched, float* loss3, float* per_triple,
               float* rows, uint32_t* row_mask, int* bad, cudaStream_t st) {
  const int threads = 256;
  ProfScope ps(PROF_BPR, st);
  k_bpr<<<(int)ceil_div(batch * 32, threads), threads, 0, st>>>(num_nodes, ld, batch, users, pos, neg, out, e0,
                                                                (float)(1.0 / (double)batch), rows, touched,
                                                                per_triple, bad);
  LGC_LAUNCH_CHECK();
  k_bpr_scatter<<<(int)ceil_div(batch * 3 * 32, threads), threads, 0, st>>>(
      (int)(batch * 3), ld, touched, rows, e0, (float)(decay / (double)batch), alpha0, grad_out, grad_e0, row_mask);
  LGC_LAUNCH_CHECK();
  k_bpr_reduce<<<1, 256, 0, st>>>(per_triple, batch, decay, bad, loss3);
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

}  // namespace lgc

using namespace lgc;

extern "C" int lgc_pair_scores(int ld, const float* out, const int64_t* pairs, int64_t n_pairs,
                               float* score, void* stream) {
  LGC_REQUIRE(out && pairs && score, "null argument");
  LGC_REQUIRE(ld > 0 && ld % 4 == 0, "ld must be a positive multiple of 4");
  if (n_pairs == 0) return LGC_OK;
  k_pair_scores<<<(int)ceil_div(n_pairs * 32, 256), 256, 0, (cudaStream_t)stream>>>(out, ld, pairs,
                                                                                     n_pairs, score);
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

extern "C" int lgc_scatter_add_rows(int64_t n, int ld, const int64_t* idx, const float* rows, float* table,
                                    void* stream) {
  LGC_REQUIRE(idx && rows && table, "null argument");
  LGC_REQUIRE(ld > 0 && ld % 4 == 0 && ld <= 128 * kMaxVecPerLane, "unsupported ld");
  LGC_REQUIRE(n >= 0 && n < (1 << 24), "row count out of range");
  if (n == 0) return LGC_OK;
  k_scatter_add_rows<<<(int)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>((int)n, ld, idx, rows, table);
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

// 16 bytes of flags | per-triple terms | touched ids (when the caller passes none) | the 3 * batch
// gradient rows (sized for the widest supported row, ld = 256)
extern "C" size_t lgc_bpr_workspace_bytes(int64_t batch) {
  return 16 + (size_t)batch * 2 * sizeof(float) + (size_t)batch * 3 * sizeof(int32_t) +
         (size_t)batch * 3 * 128 * kMaxVecPerLane * sizeof(float) + 64;
}

extern "C" int lgc_bpr_loss_grad(int64_t num_nodes, int ld, int64_t batch, const int64_t* users,
                                 const int64_t* pos, const int64_t* neg, const float* out,
                                 const float* e0, double decay, float alpha0, float* grad_out,
                                 float* grad_e0, int32_t* touched, float* loss3, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  LGC_REQUIRE(users && pos && neg && out && e0 && grad_out && loss3 && workspace, "null argument");
  LGC_REQUIRE(batch > 0, "batch must be positive");
  LGC_REQUIRE(ld > 0 && ld % 4 == 0 && ld <= 128 * kMaxVecPerLane, "unsupported ld");
  if (workspace_bytes < lgc_bpr_workspace_bytes(batch)) {
    set_error("lgc_bpr_loss_grad: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int* bad = (int*)workspace;
  float* per_triple = (float*)((char*)workspace + 16);
  int32_t* own_touched = (int32_t*)(per_triple + batch * 2);
  float* rows = (float*)(((uintptr_t)(own_touched + batch * 3) + 63) / 64 * 64);
  LGC_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  int rc = bpr_launch(num_nodes, ld, batch, users, pos, neg, out, e0, decay, alpha0, grad_out, grad_e0,
                      touched ? touched : own_touched, loss3, per_triple, rows, nullptr, bad, st);
  return rc;
}

static int adam_launch(int64_t n, float* p, const float* g, float* m, float* v, AdamScalars s,
                       const AdamScalars* s_dev, cudaStream_t st) {
  LGC_REQUIRE(p && g && m && v, "null argument");
  LGC_REQUIRE(n >= 0, "bad size");
  LGC_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0,
              "buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  if (n4) {
    int grid = (int)std::min<int64_t>(ceil_div(n4, 256), kNumSMs * 16);
    k_adam<<<grid, 256, 0, st>>>(n4, (float4*)p, (const float4*)g, (float4*)m, (float4*)v, s, s_dev);
    LGC_LAUNCH_CHECK();
  }
  if (n % 4) {
    k_adam_tail<<<1, 32, 0, st>>>(n4 * 4, n, p, g, m, v, s, s_dev);
    LGC_LAUNCH_CHECK();
  }
  return LGC_OK;
}

extern "C" int lgc_adam_step(int64_t n, float* p, const float* g, float* m, float* v, double lr,
                             double beta1, double beta2, double eps, int64_t step, void* stream) {
  LGC_REQUIRE(step >= 1, "bad step");
  return adam_launch(n, p, g, m, v, make_adam_scalars(lr, beta1, beta2, eps, step), nullptr, (cudaStream_t)stream);
}

extern "C" int lgc_adam_scalars(double lr, double beta1, double beta2, double eps, int64_t step, float* h_out6) {
  LGC_REQUIRE(h_out6 && step >= 1, "bad argument");
  const AdamScalars s = make_adam_scalars(lr, beta1, beta2, eps, step);
  h_out6[0] = s.one_minus_beta1; h_out6[1] = s.beta2; h_out6[2] = s.one_minus_beta2;
  h_out6[3] = s.bc2_sqrt; h_out6[4] = s.eps; h_out6[5] = s.neg_step_size;
  return LGC_OK;
}

extern "C" int lgc_adam_step_dev(int64_t n, float* p, const float* g, float* m, float* v,
                                 const float* d_scalars6, void* stream) {
  LGC_REQUIRE(d_scalars6, "null scalars");
  return adam_launch(n, p, g, m, v, AdamScalars{}, reinterpret_cast<const AdamScalars*>(d_scalars6),
                     (cudaStream_t)stream);
}
