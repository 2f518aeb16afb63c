// Shared helpers for liblgc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "lgc_b200.h"

namespace lgc {

void set_error(const std::string& msg);

#define LGC_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::lgc::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                       __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
      return LGC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define LGC_REQUIRE(cond, msg)                                                           \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::lgc::set_error(std::string(msg) + " [" #cond "]");                               \
      return LGC_ERR_INVALID;                                                            \
    }                                                                                    \
  } while (0)

extern long long g_launch_count;   // kernels launched by this library (claim for bench.py)
#define LGC_LAUNCH_CHECK()              \
  do {                                  \
    ++::lgc::g_launch_count;            \
    LGC_CUDA(cudaGetLastError());       \
  } while (0)

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py only).
enum ProfTag : int {
  PROF_LIGHT = 0,    // + EpiMode
  PROF_HEAVY = 4,    // + EpiMode
  PROF_FINISH = 8,   // + EpiMode
  PROF_BPR = 12,
  PROF_MISC = 13,
  PROF_EXCHANGE = 14,   // lgc_item_exchange (includes the wait for the slowest rank)
  PROF_SCORE_CONVERT = 16,
  PROF_SCORE_GEMM = 17,
  PROF_SCORE_SELECT = 18,
  PROF_SCORE_RESCORE = 19,
  PROF_SCORE_FINAL = 20,
  PROF_SCORE_EXHAUSTIVE = 21,
  PROF_SCORE_SCAN = 22,
  PROF_NUM_TAGS = 24
};
bool prof_enabled();
void prof_record(int tag, cudaStream_t st, bool begin);
struct ProfScope {
  int tag; cudaStream_t st; bool on;
  ProfScope(int t, cudaStream_t s) : tag(t), st(s), on(prof_enabled()) { if (on) prof_record(tag, st, true); }
  ~ProfScope() { if (on) prof_record(tag, st, false); }
};

// NVTX ranges around the C-ABI entry points (LGC_NVTX=1): what an nsys / ncu --nvtx timeline of a
// training or scoring run is read by (the reference has no tracing at all).
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtx_push(name); }
  ~NvtxRange() { nvtx_pop(); }
};

constexpr int kNumSMs = 148;  // B200 (grid sizing of the small helper kernels)
constexpr int kMaxDevices = 64;

// Device of the calling thread, clamped into [0, kMaxDevices): index of the per-device caches of
// function attributes / occupancy (cudaFuncSetAttribute is per device; a process may use several).
static inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}
// SM count of the current device (cached per device)
static inline int device_sm_count() {
  static int sms[kMaxDevices] = {};
  const int slot = current_device_slot();
  if (!sms[slot]) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    sms[slot] = n;
  }
  return sms[slot];
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float4 ld_f4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// read-only 128-bit load as plain PTX (inside hot loops __ldg made ptxas rebuild a memory descriptor,
// two R2UR, for every load)
__device__ __forceinline__ float4 ldg_f4_ptx(const float* p) {
  float4 v;
  asm("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_f4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
// streaming store: written once, not re-read by this kernel
__device__ __forceinline__ void st_f4_cs(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ float4 ld_f4_cs(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 fma4(float w, float4 x, float4 a) {
  a.x = fmaf(w, x.x, a.x);
  a.y = fmaf(w, x.y, a.y);
  a.z = fmaf(w, x.z, a.z);
  a.w = fmaf(w, x.w, a.w);
  return a;
}
// w * x + a with two packed FFMA2 (fma.rn.f32x2, sm_100): same roundings as four fmaf, half the
// issue slots.
__device__ __forceinline__ float4 fma4_packed(float w, float4 x, float4 a) {
  unsigned long long ww, x0, x1, a0, a1, r0, r1;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "f"(x.z), "f"(x.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(a.z), "f"(a.w));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r0) : "l"(ww), "l"(x0), "l"(a0));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r1) : "l"(ww), "l"(x1), "l"(a1));
  float4 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(r0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(r1));
  return r;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ---- table-row geometry: a row of `ld` floats = VEC float4 = L lanes x V float4 per lane.
struct RowShape {
  int L, V;
};
static inline bool row_shape(int ld, RowShape* rs) {
  if (ld <= 0 || ld % 4) return false;
  int vec = ld / 4, L = 16;
  while (L > 1 && vec % L) L >>= 1;
  int V = vec / L;
  if (V > 5) return false;
  rs->L = L;
  rs->V = V;
  return true;
}

}  // namespace lgc

// Sweep schedule of the high-degree rows (sweep.cu): built lazily per accumulator-slot count.
namespace lgc {
struct SweepSched;
void sweep_destroy(SweepSched* s);
}
constexpr int kMaxSweepScheds = 4;

// Opaque graph handle (definition shared by the .cu files).
struct lgc_graph {
  int64_t num_nodes = 0, nnz = 0;
  int is_symmetric = 0;
  int light_max_degree = 0;
  int64_t num_heavy_rows = 0, num_chunks = 0, num_split_rows = 0;
  int32_t* rowptr = nullptr;
  int32_t* src = nullptr;
  int32_t* eid = nullptr;
  float* w_hat = nullptr;
  float* deg = nullptr;
  float* dis = nullptr;
  // heavy-row schedule
  // (source, weight) arrays the heavy-row kernel gathers through: equal to src / w_hat except that
  // the entries of split (hub) rows are sorted by source, so that a hub's chunks are contiguous
  // source ranges and the chunk list can be ordered by source for L2 reuse
  int32_t* hsrc = nullptr;
  float* hw = nullptr;
  int4* chunks = nullptr;       // {row, beg, end, partial_slot or -1}, ordered by first source
  int4* split_rows = nullptr;   // {row, first_slot, n_slots, 0}
  int64_t num_partial_slots = 0;   // one [ld] partial row per chunk of a split row (workspace)
  int64_t num_cols = 0;            // rows of the gathered table (== num_nodes unless rectangular)
  // lazily built sweep schedules (mutable cache: the handle is logically const for its users)
  mutable lgc::SweepSched* sweep[kMaxSweepScheds] = {};
  mutable int sweep_failed[kMaxSweepScheds] = {};   // slot counts for which no schedule exists
  mutable int n_sweep_failed = 0;
};
