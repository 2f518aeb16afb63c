// Device-side BPR triple sampler: the semantics of the reference's
// `batch_loader` (src/utils_v2.py:168-181) without pandas, the Python lambdas or the host->device
// copies that make it cost ~24 ms per 1024-batch next to a ~4 ms training step:
//   users  `random.sample(purchasers, batch)`   distinct, uniform without replacement
//   pos    `random.choice(user's train purchases)`
//   neg    `randint(0, n_items-1) + n_users` until it is outside the user's ignore list
// The reference is unseeded (`config.yaml:2` random_seed is never read), so parity is
// distributional; here every draw is a pure function of (seed, step, triple, attempt), which makes
// a batch reproducible.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace lgc {
namespace {

__device__ __forceinline__ unsigned long long mix64s(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
// counter-based generator: 64 random bits per (seed, step, lane, counter)
__device__ __forceinline__ unsigned long long rng64(unsigned long long seed, unsigned long long step,
                                                    unsigned long long lane, unsigned long long ctr) {
  return mix64s(mix64s(seed ^ 0x9e3779b97f4a7c15ULL * (step + 1)) ^ mix64s(lane * 0xd1342543de82ef95ULL + ctr));
}
// unbiased enough for sampling: floor(r * n / 2^64)
__device__ __forceinline__ long long bounded(unsigned long long r, long long n) {
  return (long long)__umul64hi(r, (unsigned long long)n);
}

// users without replacement, reproducibly: every purchaser gets a random 64-bit key, the `batch`
// smallest keys win (radix sort of the keys; their order is the batch order)
__global__ void k_sample_keys(long long n_purchasers, unsigned long long seed, unsigned long long step,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_purchasers) return;
  keys[i] = rng64(seed, step, (unsigned long long)i, 0x5eedULL);
  vals[i] = (int)i;
}

__global__ void k_sample_triples(const int* __restrict__ chosen, const int64_t* __restrict__ purchasers,
                                 const int64_t* __restrict__ pos_ptr, const int64_t* __restrict__ pos_items,
                                 const int64_t* __restrict__ ign_ptr, const int64_t* __restrict__ ign_items,
                                 long long n_users, long long n_items, long long batch, unsigned long long seed,
                                 unsigned long long step, int64_t* __restrict__ users, int64_t* __restrict__ pos,
                                 int64_t* __restrict__ neg) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= batch) return;
  unsigned long long ctr = 0;
  const long long p = chosen[t];
  users[t] = purchasers[p];
  // ---- positive: uniform over the user's train purchases
  const long long pb = pos_ptr[p], pn = pos_ptr[p + 1] - pb;
  pos[t] = pos_items[pb + bounded(rng64(seed, step, (unsigned long long)t, ctr++), pn)];
  // ---- negative: rejection outside the (sorted) ignore list
  const long long ib = ign_ptr[p], in = ign_ptr[p + 1] - ib;
  for (int attempt = 0;; ++attempt) {
    const long long cand = bounded(rng64(seed, step, (unsigned long long)t, ctr++), n_items) + n_users;
    long long lo = 0, hi = in;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (ign_items[ib + mid] < cand) lo = mid + 1; else hi = mid;
    }
    if (!(lo < in && ign_items[ib + lo] == cand) || attempt > 4096) { neg[t] = cand; break; }
  }
}

struct SampleLayout {
  size_t off_keys_in, off_keys_out, off_vals_in, off_vals_out, off_tmp, tmp_bytes, bytes;
};
SampleLayout sample_layout(long long n_purchasers) {
  SampleLayout L;
  const size_t n = (size_t)std::max<long long>(n_purchasers, 1);
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += (b + 255) / 256 * 256; return o; };
  L.off_keys_in = take(n * 8); L.off_keys_out = take(n * 8);
  L.off_vals_in = take(n * 4); L.off_vals_out = take(n * 4);
  L.tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.tmp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, (int)n);
  L.off_tmp = take(L.tmp_bytes);
  L.bytes = off + 256;
  return L;
}

}  // namespace
}  // namespace lgc

using namespace lgc;

extern "C" size_t lgc_sample_triples_workspace_bytes(int64_t n_purchasers, int64_t batch) {
  (void)batch;
  return sample_layout(n_purchasers).bytes;
}

extern "C" int lgc_sample_triples(int64_t n_purchasers, const int64_t* purchasers, const int64_t* pos_ptr,
                                  const int64_t* pos_items, const int64_t* ign_ptr, const int64_t* ign_items,
                                  int64_t n_users, int64_t n_items, int64_t batch, uint64_t seed, uint64_t step,
                                  int64_t* users, int64_t* pos, int64_t* neg, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  LGC_REQUIRE(purchasers && pos_ptr && pos_items && ign_ptr && ign_items && users && pos && neg && workspace,
              "null argument");
  LGC_REQUIRE(batch > 0 && n_items > 0 && n_users >= 0 && n_purchasers < (1LL << 31), "bad sizes");
  if (batch > n_purchasers) {
    set_error("Sample larger than population");          // random.sample's own error (ValueError)
    return LGC_ERR_INVALID;
  }
  const SampleLayout L = sample_layout(n_purchasers);
  if (workspace_bytes < L.bytes) {
    set_error("lgc_sample_triples: workspace too small");
    return LGC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  unsigned long long* keys_in = (unsigned long long*)(ws + L.off_keys_in);
  unsigned long long* keys_out = (unsigned long long*)(ws + L.off_keys_out);
  int* vals_in = (int*)(ws + L.off_vals_in);
  int* vals_out = (int*)(ws + L.off_vals_out);
  k_sample_keys<<<(int)ceil_div(n_purchasers, 256), 256, 0, st>>>(n_purchasers, seed, step, keys_in, vals_in);
  LGC_LAUNCH_CHECK();
  size_t tmp_bytes = L.tmp_bytes;
  LGC_CUDA(cub::DeviceRadixSort::SortPairs(ws + L.off_tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out,
                                           (int)n_purchasers, 0, 64, st));
  k_sample_triples<<<(int)ceil_div(batch, 256), 256, 0, st>>>(vals_out, purchasers, pos_ptr, pos_items, ign_ptr,
                                                            ign_items, n_users, n_items, batch, seed, step, users,
                                                            pos, neg);
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

// ------------------------------------------------------------------ MARK_MAPK on the device
// precision@k / recall@k of reference `LightGCN.MARK_MAPK` (src/lightgcn.py:184-189) for every
// evaluated user at once: overlap = |set(held-out items) & set(top-k)|, recall = overlap /
// len(held-out list), precision = overlap / k; then the two means. One thread per user (lists are
// short), means reduced in double precision in a fixed order.
namespace lgc {
namespace {
__global__ void k_mark_mapk(long long n_users, int k, const int64_t* __restrict__ topk,
                            const int64_t* __restrict__ held_ptr, const int64_t* __restrict__ held_items,
                            float* __restrict__ per_user) {
  const long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  const long long hb = held_ptr[u], hn = held_ptr[u + 1] - hb;
  int overlap = 0;
  for (int j = 0; j < k; ++j) {
    const int64_t it = topk[u * k + j];
    bool dup = false;                                   // set semantics on the top-k side
    for (int q = 0; q < j; ++q) dup |= (topk[u * k + q] == it);
    if (dup) continue;
    bool hit = false;
    for (long long q = 0; q < hn; ++q) hit |= (held_items[hb + q] == it);
    overlap += hit ? 1 : 0;
  }
  per_user[2 * u] = (float)overlap / (float)k;
  per_user[2 * u + 1] = hn > 0 ? (float)overlap / (float)hn : 0.f;
}
__global__ void k_mean2(const float* __restrict__ per_user, long long n, double* __restrict__ out2) {
  __shared__ double s_a[256], s_b[256];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) { a += per_user[2 * i]; b += per_user[2 * i + 1]; }
  s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = n ? s_a[0] / (double)n : 0.0; out2[1] = n ? s_b[0] / (double)n : 0.0; }
}
}  // namespace
}  // namespace lgc

extern "C" int lgc_mark_mapk(int64_t n_users, int k, const int64_t* topk_items, const int64_t* held_ptr,
                             const int64_t* held_items, float* per_user, double* out2, void* stream) {
  LGC_REQUIRE(topk_items && held_ptr && held_items && per_user && out2, "null argument");
  LGC_REQUIRE(n_users >= 0 && k >= 1, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_users > 0) {
    k_mark_mapk<<<(int)ceil_div(n_users, 256), 256, 0, st>>>(n_users, k, topk_items, held_ptr, held_items, per_user);
    LGC_LAUNCH_CHECK();
  }
  k_mean2<<<1, 256, 0, st>>>(per_user, n_users, out2);
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}
