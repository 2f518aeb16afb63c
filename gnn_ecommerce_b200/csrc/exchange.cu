// Item-row exchange of the multi-GPU step as ONE kernel over peer memory (NVLink / NVSwitch P2P):
// reduce-scatter of the ranks' partial item sums, the fused LGConv epilogue on the reduced slice and
// the all-gather of the results, with no NCCL call and no host synchronisation.
//
// What it replaces: in the bipartite-sharded step (gnn_ecommerce_b200/sharded.py)
// every rank holds the partial sums of (A_hat x)[items] over its own users. The reference has no
// multi-GPU path at all (src/train_lightgcn.py:35-37 moves one graph to one device); the sum over
// ranks completes the `scatter_add_` of PyG's LGConv aggregate (call site src/lightgcn.py:96) for
// the item rows, and the epilogue is what the reference runs as separate ATen passes afterwards
// (running layer mean src/lightgcn.py:93,97; backward Horner add; torch.optim.Adam,
// src/train_lightgcn.py:147). Before this kernel the step issued ncclAllReduce on the [n_items, ld]
// partials and then lgc_epilogue_apply on every rank.
//
// Memory: every rank allocates one ARENA (lgc_peer_arena_alloc: cudaMalloc + CUDA IPC handle) that
// holds, at the SAME offsets on every rank, the partial-sum tables, every replicated item table the
// epilogues write (x_l, out, E0, m, v) and a 256-byte control block; the peers' arenas are mapped
// with lgc_peer_arena_open. One launch on every rank then does
//   A  barrier: block 0 publishes "my partials are complete" into every peer's control block
//      (st.release.sys), every block waits for all peers' tickets (ld.acquire.sys);
//   B  rank r owns the float4 range [v_beg, v_end) of the table: it loads that range from ALL
//      arenas (P2P loads over NVLink, fixed rank order 0..world-1: deterministic), applies the
//      epilogue with the replicated local operands and stores the results into EVERY arena
//      (P2P stores): all replicas stay bit-identical because one rank computes each element. ADAM
//      sends only the new weights; the moments m, v of a row stay with the rank that owns the row;
//   C  the last block of the rank to finish publishes "my stores are performed" and waits for the
//      same ticket of every peer; the kernel ends when every arena holds every slice.
// Tickets are monotone (2 per launch, counted in the control block, so a captured CUDA graph can
// replay the launch) and every wait is bounded: a timeout raises the error word of the control block
// instead of hanging the GPU (read by lgc_peer_exchange_status).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "epilogue.cuh"

namespace lgc {
namespace {

struct PeerCtrl {              // one per arena, zero-initialised by lgc_peer_arena_alloc
  unsigned int flag[16];       // flag[q]: latest ticket rank q has published to the owner of this block
  unsigned int epoch;          // exchanges this rank has completed (ticket base = 2 * epoch)
  unsigned int done;           // blocks of the running launch whose stores are performed
  unsigned int error;          // != 0: a wait timed out (1 = phase A, 2 = phase C)
  unsigned int pad[45];
};
static_assert(sizeof(PeerCtrl) == 256, "control block layout");

struct XArgs {
  int world, rank;
  char* base[LGC_PEER_MAX];    // arena of every rank as mapped in this process
  size_t ctrl_off;             // byte offset of the control block
  size_t part_off;             // byte offset of the partial-sum table
  size_t out_off[3];           // byte offsets of the tables the epilogue writes
  int64_t v_beg, v_end;        // float4 range of the table this rank reduces
  long long timeout_cycles;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data: read once per launch, must not come from a stale L1 line
__device__ __forceinline__ void ld_peer(const float* p, float (&r)[4]) {
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_peer(float* p, const float (&r)[4]) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]),
               "f"(r[3]) : "memory");
}

// wait until rank q's ticket in my control block reaches `ticket` (wrap-safe compare); bounded
__device__ __forceinline__ bool wait_ticket(const unsigned int* flag, unsigned int ticket, long long timeout) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - ticket) < 0) {
    if (clock64() - t0 > timeout) return false;
    __nanosleep(64);
  }
  return true;
}

// The values the epilogue stores, in the rounding sequence of epilogue.cuh / spmm.cu (bit-identical to
// lgc_epilogue_apply on the same sums): PLAIN -> {y}, ADAM -> {p, m, v}, FWD_FINAL -> {acc}.
template <int MODE>
__device__ __forceinline__ void epi_values(const EpiArgs& a, const float (&s)[4], const EpiPre<MODE, 4>& p,
                                           float (&out)[3][4]) {
  if constexpr (MODE == EPI_PLAIN) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[0][i] = a.scale * s[i];
    if (a.addend) {
#pragma unroll
      for (int i = 0; i < 4; ++i) out[0][i] = fmaf(a.beta, p.q[0][i], out[0][i]);
    }
  } else if constexpr (MODE == EPI_ADAM) {
    const AdamScalars ad = a.adam_dev ? *a.adam_dev : a.adam;
    const float ib = __frcp_rn(ad.bc2_sqrt);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[0][i] = p.q[1][i]; out[1][i] = p.q[2][i]; out[2][i] = p.q[3][i];
      adam_update_fast(out[0][i], out[1][i], out[2][i], fmaf(a.scale, s[i], p.q[0][i]), ad, ib);
    }
  } else {  // EPI_FWD_FINAL
#pragma unroll
    for (int i = 0; i < 4; ++i) out[0][i] = __fmul_rn(p.q[0][i], a.ah[0]);
#pragma unroll
    for (int h = 1; h < kMaxHist; ++h)
      if (h < a.n_hist) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[0][i] = __fadd_rn(out[0][i], __fmul_rn(p.q[h][i], a.ah[h]));
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[0][i] = __fadd_rn(out[0][i], __fmul_rn(s[i], a.a1));
  }
}

// WMAX = upper bound of the world size (2, 4 or 8: sizes the register tile of peer loads), U = elements
// per thread and pass. Remote loads take ~2-3 us over NVLink, so a pass keeps U * world 16-byte loads
// in flight per thread and the grid is sized so that a slice needs one or two passes.
template <int MODE, int WMAX, int U>
__global__ void __launch_bounds__(256, 2)
k_item_exchange(XArgs x, EpiArgs a) {
  PeerCtrl* my = reinterpret_cast<PeerCtrl*>(x.base[x.rank] + x.ctrl_off);
  __shared__ unsigned int s_ticket;
  __shared__ int s_last;
  // `epoch` changes only at the very end of a launch (last block, after every block has counted itself
  // into `done`), so all blocks of a launch read the same value
  if (threadIdx.x == 0) s_ticket = 2u * *reinterpret_cast<volatile unsigned int*>(&my->epoch);
  __syncthreads();
  const unsigned int t_ready = s_ticket + 1u, t_done = s_ticket + 2u;

  // ---- A: every rank's partial sums are complete (they were written by the kernel before this one)
  if ((int)threadIdx.x < x.world) {
    const int q = threadIdx.x;
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(&reinterpret_cast<PeerCtrl*>(x.base[q] + x.ctrl_off)->flag[x.rank], t_ready);
    }
    if (!wait_ticket(&my->flag[q], t_ready, x.timeout_cycles)) my->error = 1;
  }
  __syncthreads();

  // ---- B: reduce my slice over the ranks, epilogue, store into every arena
  const int64_t n_thr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = x.v_beg + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < x.v_end; i0 += n_thr * U) {
    EpiPre<MODE, 4> pre[U];
    float part[U][WMAX][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * n_thr;
      if (i < x.v_end) {
        const size_t off = (size_t)i * 4;
#pragma unroll
        for (int q = 0; q < WMAX; ++q)
          if (q < x.world) ld_peer(reinterpret_cast<const float*>(x.base[q] + x.part_off) + off, part[u][q]);
        epi_preload_w<MODE, 4>(a, off, pre[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * n_thr;
      if (i < x.v_end) {
        const size_t off = (size_t)i * 4;
        float s[4] = {part[u][0][0], part[u][0][1], part[u][0][2], part[u][0][3]};
#pragma unroll
        for (int q = 1; q < WMAX; ++q)
          if (q < x.world) {
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] = __fadd_rn(s[c], part[u][q][c]);
          }
        float out[3][4];
        epi_values<MODE>(a, s, pre[u], out);
#pragma unroll
        for (int q = 0; q < WMAX; ++q)
          if (q < x.world) st_peer(reinterpret_cast<float*>(x.base[q] + x.out_off[0]) + off, out[0]);
        if constexpr (MODE == EPI_ADAM) {      // the moments of a row are only ever read by the rank that owns it
          stv_stream<4>(a.m + off, out[1]);
          stv_stream<4>(a.v + off, out[2]);
        }
      }
    }
  }

  // ---- C: all stores of this rank are performed -> tell the peers; leave when all peers have told me
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&my->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if ((int)threadIdx.x < x.world) {
    const int q = threadIdx.x;
    __threadfence_system();
    st_release_sys(&reinterpret_cast<PeerCtrl*>(x.base[q] + x.ctrl_off)->flag[x.rank], t_done);
    if (!wait_ticket(&my->flag[q], t_done, x.timeout_cycles)) my->error = 2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    my->done = 0;
    my->epoch = (s_ticket >> 1) + 1u;
    __threadfence();
  }
}

bool inside(const void* p, const char* base, size_t bytes, size_t need) {
  const char* c = static_cast<const char*>(p);
  return c >= base && c + need <= base + bytes;
}

}  // namespace
}  // namespace lgc

using namespace lgc;

static_assert(sizeof(cudaIpcMemHandle_t) == LGC_PEER_HANDLE_BYTES, "CUDA IPC handle size");

extern "C" int lgc_peer_arena_alloc(size_t bytes, void** d_base, void* h_handle) {
  LGC_REQUIRE(d_base && h_handle && bytes >= LGC_PEER_CTRL_BYTES, "bad argument");
  void* p = nullptr;
  LGC_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error(std::string("lgc_peer_arena_alloc: ") + cudaGetErrorString(e));
    return LGC_ERR_CUDA;
  }
  memcpy(h_handle, &h, sizeof(h));
  *d_base = p;
  return LGC_OK;
}

extern "C" int lgc_peer_arena_open(const void* h_handle, void** d_peer_base) {
  LGC_REQUIRE(h_handle && d_peer_base, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle, sizeof(h));
  void* p = nullptr;
  LGC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_peer_base = p;
  return LGC_OK;
}

extern "C" int lgc_peer_arena_close(void* d_peer_base) {
  if (d_peer_base) LGC_CUDA(cudaIpcCloseMemHandle(d_peer_base));
  return LGC_OK;
}

extern "C" int lgc_peer_arena_free(void* d_base) {
  if (d_base) LGC_CUDA(cudaFree(d_base));
  return LGC_OK;
}

extern "C" int lgc_item_exchange(const lgc_peer_exchange* px, const lgc_spmm_epilogue* e, void* stream) {
  LGC_REQUIRE(px && e, "null argument");
  LGC_REQUIRE(px->world >= 1 && px->world <= LGC_PEER_MAX && px->rank >= 0 && px->rank < px->world, "bad world / rank");
  LGC_REQUIRE(px->n_rows >= 0 && px->ld > 0 && px->ld % 4 == 0, "bad table shape");
  LGC_REQUIRE(e->mode == LGC_EPI_PLAIN || e->mode == LGC_EPI_ADAM || e->mode == LGC_EPI_FWD_FINAL,
              "lgc_item_exchange supports the PLAIN, ADAM and FWD_FINAL epilogues");
  const char* base = static_cast<const char*>(px->bases[px->rank]);
  const size_t table_bytes = (size_t)px->n_rows * px->ld * sizeof(float);
  for (int q = 0; q < px->world; ++q) LGC_REQUIRE(px->bases[q], "null arena");
  LGC_REQUIRE(px->ctrl_off % 256 == 0 && px->ctrl_off + LGC_PEER_CTRL_BYTES <= px->arena_bytes, "control block outside the arena");
  LGC_REQUIRE(inside(px->part, base, px->arena_bytes, table_bytes), "partial sums outside the arena");

  XArgs x;
  x.world = px->world; x.rank = px->rank;
  for (int q = 0; q < LGC_PEER_MAX; ++q) x.base[q] = q < px->world ? static_cast<char*>(px->bases[q]) : nullptr;
  x.ctrl_off = px->ctrl_off;
  x.part_off = (size_t)(reinterpret_cast<const char*>(px->part) - base);
  EpiArgs a;
  a.addend = e->addend; a.a1 = e->a1; a.scale = e->scale; a.beta = e->beta;
  float* outs[3] = {nullptr, nullptr, nullptr};
  int n_out = 1;
  switch (e->mode) {
    case LGC_EPI_PLAIN: outs[0] = e->y; break;
    case LGC_EPI_ADAM:
      LGC_REQUIRE(e->addend && e->p && e->m && e->v && (e->step >= 1 || e->adam_scalars),
                  "ADAM needs addend, p, m, v and step >= 1 (or device scalars)");
      outs[0] = e->p;                     // m, v: local stores (rows this rank owns)
      a.p = e->p; a.m = e->m; a.v = e->v;
      if (e->adam_scalars) a.adam_dev = reinterpret_cast<const AdamScalars*>(e->adam_scalars);
      else a.adam = make_adam_scalars(e->lr, e->beta1, e->beta2, e->eps, e->step);
      break;
    default:
      LGC_REQUIRE(e->acc && e->n_hist >= 1 && e->n_hist <= kMaxHist, "FWD_FINAL needs acc and 1..6 layer tables");
      outs[0] = e->acc;
      a.n_hist = e->n_hist;
      for (int i = 0; i < e->n_hist; ++i) {
        LGC_REQUIRE(e->hist[i], "FWD_FINAL: null layer table");
        a.hist[i] = e->hist[i];
        a.ah[i] = e->ah[i];
      }
  }
  for (int o = 0; o < 3; ++o) {
    x.out_off[o] = 0;
    if (o < n_out) {
      LGC_REQUIRE(inside(outs[o], base, px->arena_bytes, table_bytes),
                  "every table the epilogue writes must lie inside the arena (same offset on every rank)");
      x.out_off[o] = (size_t)(reinterpret_cast<const char*>(outs[o]) - base);
    }
  }
  // contiguous float4 slices, whole rows per rank
  const int64_t vec = px->ld / 4;
  const int64_t rows_per = ceil_div(px->n_rows, px->world);
  x.v_beg = std::min<int64_t>(px->n_rows, rows_per * px->rank) * vec;
  x.v_end = std::min<int64_t>(px->n_rows, rows_per * (px->rank + 1)) * vec;
  x.timeout_cycles = px->timeout_ms > 0 ? (long long)px->timeout_ms * 2000000LL : 40000000000LL;   // ~2 GHz

  const int64_t n_vec = x.v_end - x.v_beg;
  cudaStream_t st = (cudaStream_t)stream;
  NvtxRange nvtx("lgc_item_exchange");
  ProfScope ps(PROF_EXCHANGE, st);
  // at most one CTA per SM: the launch overlaps the rows kernel of the same layer and spins in phase A
  // until the slowest rank arrives, so it must leave most of every SM to that kernel; U * world loads in
  // flight per thread (LGC_XCHG_CTAS: development override of the CTA cap)
  static const int cta_cap = [] { const char* e = getenv("LGC_XCHG_CTAS"); return e ? std::max(1, atoi(e)) : 0; }();
#define LGC_XCASE(WMAX, U)                                                                                     \
  {                                                                                                            \
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_vec, 256 * (U)), cta_cap ? cta_cap : device_sm_count())); \
    switch (e->mode) {                                                                                         \
      case LGC_EPI_PLAIN: k_item_exchange<EPI_PLAIN, WMAX, U><<<grid, 256, 0, st>>>(x, a); break;              \
      case LGC_EPI_ADAM: k_item_exchange<EPI_ADAM, WMAX, U><<<grid, 256, 0, st>>>(x, a); break;                \
      default: k_item_exchange<EPI_FWD_FINAL, WMAX, U><<<grid, 256, 0, st>>>(x, a); break;                     \
    }                                                                                                          \
  }
  if (px->world <= 2) LGC_XCASE(2, 4)
  else if (px->world <= 4) LGC_XCASE(4, 2)
  else LGC_XCASE(8, 1)
#undef LGC_XCASE
  LGC_LAUNCH_CHECK();
  return LGC_OK;
}

extern "C" int lgc_peer_exchange_status(const lgc_peer_exchange* px, int32_t* h_error, int64_t* h_epoch) {
  LGC_REQUIRE(px && px->rank >= 0 && px->rank < LGC_PEER_MAX && px->bases[px->rank], "bad argument");
  PeerCtrl c;
  LGC_CUDA(cudaMemcpy(&c, static_cast<const char*>(px->bases[px->rank]) + px->ctrl_off, sizeof(c),
                      cudaMemcpyDeviceToHost));
  if (h_error) *h_error = (int32_t)c.error;
  if (h_epoch) *h_epoch = (int64_t)c.epoch;
  return LGC_OK;
}
