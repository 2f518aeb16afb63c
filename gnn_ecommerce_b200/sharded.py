"""Row-partitioned multi-GPU BPR training step (one process per GPU, `torch.distributed`/NCCL).

The single-GPU step (`trainer.FusedBPRTrainer`, reference `src/train_lightgcn.py:129-151`) is
sharded the way BASELINE.json's north_star states it: destination rows are partitioned into
contiguous ranges balanced by work, every rank owns its rows of the embedding table, the Adam
moments and all per-layer tables, and each LGConv layer is

    local fused SpMM over the rank's rows (reads the all-gathered table)  ->  all-gather of the shards

The loss needs the final embeddings of <= 3*batch nodes: every rank contributes the rows it owns
to one small all-reduce (the "BPR sparse rows" exchange), evaluates the tiny loss redundantly with
the same kernel as the single-GPU path, and scatters the gradient rows locally -- the dense
gradient table is never communicated. Layout trick: shards are padded to a common row count
`max_rows` and sources are renumbered to `owner * max_rows + local`, so the all-gather output IS
the gather table of the next layer (no re-packing, equal-sized NCCL chunks).

The arithmetic is the same sm_100a kernels (`lgc_spmm_ex`, `lgc_bpr_loss_grad`) through the C ABI;
a `backend` object carries those calls so the orchestration can be exercised on CPU with `gloo`
(tests inject a checker backend; the product backend is `CudaBackend` and has no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

ROW_COST = 4.0   # epilogue streams of one row cost about as much as gathering four neighbours


# --------------------------------------------------------------------------------- partition
class RowPartition:
    """Contiguous destination-row ranges with balanced cost (in-degree + ROW_COST per row)."""

    def __init__(self, degree: np.ndarray, world: int):
        degree = np.asarray(degree, dtype=np.float64)
        n = degree.shape[0]
        cost = np.cumsum(degree + ROW_COST)
        total = cost[-1] if n else 0.0
        bounds = [0]
        for r in range(1, world):
            bounds.append(int(np.searchsorted(cost, total * r / world, side="left")))
        bounds.append(n)
        self.bounds = np.maximum.accumulate(np.asarray(bounds, dtype=np.int64))
        self.world, self.num_nodes = world, n
        self.max_rows = int(max(1, np.diff(self.bounds).max()))
        self.max_rows = (self.max_rows + 3) // 4 * 4

    def lo(self, rank: int) -> int:
        return int(self.bounds[rank])

    def hi(self, rank: int) -> int:
        return int(self.bounds[rank + 1])

    def owner(self, ids: Tensor) -> Tensor:
        b = torch.as_tensor(self.bounds[1:], device=ids.device)
        return torch.bucketize(ids, b, right=True)

    def padded_id(self, ids: Tensor) -> Tensor:
        """Global node id -> row of the all-gathered `[world * max_rows, ld]` table."""
        own = self.owner(ids)
        lo = torch.as_tensor(self.bounds[:-1], device=ids.device)[own]
        return own * self.max_rows + (ids - lo)

    def unpad(self, table: Tensor) -> Tensor:
        """`[world * max_rows, ...]` -> `[num_nodes, ...]` (drops the padding rows)."""
        parts = [table[r * self.max_rows: r * self.max_rows + self.hi(r) - self.lo(r)] for r in range(self.world)]
        return torch.cat(parts, 0)


# --------------------------------------------------------------------------------- CUDA backend
class CudaBackend:
    """The product backend: raw pointers into liblgc_b200.so. No CPU path."""

    def __init__(self):
        from . import _capi
        self._capi, self._lib = _capi, _capi.lib()

    def global_w_hat(self, edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int):
        """gcn_norm weights of the GLOBAL graph in edge-list order + the in-degree counts."""
        from .graph import Graph
        g = Graph(edge_index, edge_weight, num_nodes)
        w = g.w_hat_edge_order()
        deg = torch.diff(g.arrays()["rowptr"].long())
        sym = g.is_symmetric
        del g
        return w, deg, sym

    def build_rect(self, src: Tensor, dst: Tensor, w: Tensor, n_rows: int, n_cols: int):
        from .graph import _ptr, _stream
        ei = torch.stack([src, dst]).contiguous()
        handle = C.c_void_p()
        with torch.cuda.device(ei.device):
            rc = self._lib.lgc_graph_build_rect(n_rows, n_cols, ei.size(1), _ptr(ei), _ptr(w.contiguous()),
                                                _stream(), C.byref(handle))
        self._capi.check(rc, "lgc_graph_build_rect")
        return handle

    def destroy(self, handle) -> None:
        self._lib.lgc_graph_destroy(handle)

    def workspace(self, handle, ld: int, device) -> Tensor:
        return torch.empty(max(256, self._lib.lgc_spmm_workspace_bytes(handle, ld)), dtype=torch.uint8, device=device)

    def spmm_ex(self, handle, ld: int, x: Tensor, ws: Tensor, mode: int, *, y=None, acc=None, xrow=None,
                addend=None, a0=0.0, a1=0.0, scale=1.0, beta=0.0, p=None, m=None, v=None, lr=0.0,
                betas=(0.9, 0.999), eps=1e-8, step=1) -> None:
        from .graph import _ptr, _stream
        e = self._capi.SpmmEpilogue(mode=mode, a0=a0, a1=a1, scale=scale, beta=beta, y=_ptr(y), acc=_ptr(acc),
                                    xrow=_ptr(xrow), addend=_ptr(addend), p=_ptr(p), m=_ptr(m), v=_ptr(v),
                                    lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, step=step)
        with torch.cuda.device(x.device):
            rc = self._lib.lgc_spmm_ex(handle, ld, _ptr(x), C.byref(e), _ptr(ws), ws.numel(), _stream())
        self._capi.check(rc, "lgc_spmm_ex")

    def bpr(self, outc: Tensor, e0c: Tensor, batch: int, decay: float, alpha0: float):
        """BPR + L2 on the compact `[3*batch, ld]` row tables (users | pos | neg)."""
        from . import ops
        dev = outc.device
        ar = torch.arange(batch, device=dev, dtype=torch.int64)
        gc, zc = torch.zeros_like(outc), torch.zeros_like(outc)
        loss3 = ops.bpr_loss_grad(outc, e0c, ar, ar + batch, ar + 2 * batch, decay, alpha0, gc, zc)
        return loss3, gc, zc


# --------------------------------------------------------------------------------- trainer
class ShardedBPRTrainer:
    """`step()` = one mini-batch of `mini_batch_loop` over `world` GPUs; same results as the
    single-GPU fused step up to fp32 summation order (row sums are identical: every row is still
    reduced by one rank in CSR order)."""

    def __init__(self, edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int, embedding_dim: int,
                 num_layers: int, init_weight: Tensor, lr: float = 0.005, betas=(0.9, 0.999), eps: float = 1e-8,
                 alpha: Optional[Sequence[float]] = None, group=None, backend=None, ld: Optional[int] = None):
        assert num_layers >= 1, "the sharded step needs at least one propagation layer"
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = backend if backend is not None else CudaBackend()
        self.dev = edge_index.device
        self.num_nodes, self.dim, self.layers = int(num_nodes), int(embedding_dim), int(num_layers)
        if ld is None:
            from .graph import padded_dim
            ld = padded_dim(self.dim)
        self.ld = ld
        self.alpha = [float(a) for a in (alpha if alpha is not None else [1.0 / (num_layers + 1)] * (num_layers + 1))]
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0

        w_hat, deg, sym = self.backend.global_w_hat(edge_index, edge_weight, self.num_nodes)
        if not sym:
            raise RuntimeError("the sharded step needs a symmetric graph (backward reuses the operator)")
        self.part = RowPartition(deg.cpu().numpy(), self.world)
        lo, hi = self.part.lo(self.rank), self.part.hi(self.rank)
        self.lo, self.hi, self.n_local, self.max_rows = lo, hi, hi - lo, self.part.max_rows
        src, dst = edge_index[0], edge_index[1]
        mine = (dst >= lo) & (dst < hi)                       # keeps edge-list order per row
        self.n_cols = self.world * self.max_rows
        self.handle = self.backend.build_rect(self.part.padded_id(src[mine]), dst[mine] - lo, w_hat[mine],
                                              max(self.n_local, 1), self.n_cols)
        self.local_nnz = int(mine.sum().item())
        del w_hat, mine
        self.ws = self.backend.workspace(self.handle, ld, self.dev)

        def table(rows):
            return torch.zeros(rows, ld, dtype=torch.float32, device=self.dev)
        self.e0, self.m, self.v = table(self.max_rows), table(self.max_rows), table(self.max_rows)
        self.e0[: self.n_local, : self.dim] = init_weight[lo:hi].to(self.dev)
        self.out, self.y, self.z = table(self.max_rows), table(self.max_rows), table(self.max_rows)
        self.full = [table(self.n_cols), table(self.n_cols)]
        self.gfull = table(self.n_cols)                       # dL/d out, padded-global layout, sparse

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h is not None:
            try:
                self.backend.destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ collectives
    def _all_gather(self, full: Tensor, shard: Tensor) -> None:
        if self.world == 1:
            full.copy_(shard)
        else:
            dist.all_gather_into_tensor(full, shard, group=self.group)

    def _all_reduce(self, t: Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, group=self.group)

    # ------------------------------------------------------------------ forward only
    def propagate(self) -> Tensor:
        """out_local = sum_l alpha_l (A_hat^l E0)[own rows] (reference `get_embedding`)."""
        b, a, K = self.backend, self.alpha, self.layers
        cur = self.full[0]
        self._all_gather(cur, self.e0)
        for l in range(1, K + 1):
            last = l == K
            b.spmm_ex(self.handle, self.ld, cur, self.ws, 1 if l == 1 else 2, y=None if last else self.y,
                      acc=self.out, xrow=self.e0, a0=a[0], a1=a[l])
            if not last:
                cur = self.full[l & 1]
                self._all_gather(cur, self.y)
        return self.out

    def gather_table(self, shard: Tensor) -> Tensor:
        """Full `[num_nodes, dim]` table from the ranks' `[max_rows, ld]` shards (every rank)."""
        full = torch.empty(self.n_cols, self.ld, dtype=torch.float32, device=self.dev)
        self._all_gather(full, shard)
        return self.part.unpad(full)[:, : self.dim]

    # ------------------------------------------------------------------ one mini-batch
    def step(self, users: Tensor, pos: Tensor, neg: Tensor, decay: float) -> Tensor:
        b, a, K, ld = self.backend, self.alpha, self.layers, self.ld
        batch = users.numel()
        self.step_count += 1
        self.propagate()

        # ---- the <= 3*batch needed rows of out / E0: owners contribute, one small all-reduce
        ids = torch.cat([users, pos, neg]).to(device=self.dev, dtype=torch.int64)
        mine = (ids >= self.lo) & (ids < self.hi)
        loc = (ids - self.lo)[mine]
        rows = torch.zeros(2, 3 * batch, ld, dtype=torch.float32, device=self.dev)
        rows[0, mine] = self.out[loc]
        rows[1, mine] = self.e0[loc]
        self._all_reduce(rows)
        loss3, gc, zc = b.bpr(rows[0], rows[1], batch, float(decay), a[0])

        # ---- gradient rows: dL/d out for everybody (gather source of the backward), Z for owners
        pid = self.part.padded_id(ids)
        self.gfull.index_add_(0, pid, gc)
        self.z.index_add_(0, loc, zc[mine])
        g_local = self.gfull[self.rank * self.max_rows: (self.rank + 1) * self.max_rows]

        # ---- backward (Horner on the symmetric operator) + Adam on the owned rows
        cur, scale = self.gfull, a[K]
        for l in range(K - 1, 0, -1):                         # h_l = alpha_l G + A h_{l+1}
            b.spmm_ex(self.handle, ld, cur, self.ws, 0, y=self.y, addend=g_local, scale=scale, beta=a[l])
            cur = self.full[l & 1]
            self._all_gather(cur, self.y)
            scale = 1.0
        b.spmm_ex(self.handle, ld, cur, self.ws, 3, addend=self.z, scale=scale, p=self.e0, m=self.m, v=self.v,
                  lr=self.lr, betas=self.betas, eps=self.eps, step=self.step_count)
        self.gfull.index_fill_(0, pid, 0.0)
        self.z.index_fill_(0, loc, 0.0)
        return loss3

    # ------------------------------------------------------------------ views for callers / tests
    def weight(self) -> Tensor:
        return self.gather_table(self.e0)

    def embedding(self) -> Tensor:
        self.propagate()
        return self.gather_table(self.out)


def shard_users(n_users: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous user range of `rank` for user-sharded scoring (no communication)."""
    per = (n_users + world - 1) // world
    return min(n_users, rank * per), min(n_users, (rank + 1) * per)
