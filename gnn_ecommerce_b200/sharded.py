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

`BipartiteShardedTrainer` (the default for the reference's user-item graphs) shards the USERS
instead, replicates the small item table and exchanges only the item rows, through a hand-written
peer-memory kernel (`PeerArena`, `lgc_item_exchange`) or NCCL -- see its docstring.

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
        self._cache = {}

    def lo(self, rank: int) -> int:
        return int(self.bounds[rank])

    def hi(self, rank: int) -> int:
        return int(self.bounds[rank + 1])

    def _dev(self, device):
        key = str(device)
        if key not in self._cache:
            self._cache[key] = (torch.as_tensor(self.bounds[1:], device=device),
                                torch.as_tensor(self.bounds[:-1], device=device))
        return self._cache[key]

    def owner(self, ids: Tensor) -> Tensor:
        return torch.bucketize(ids, self._dev(ids.device)[0], right=True)

    def padded_id(self, ids: Tensor) -> Tensor:
        """Global node id -> row of the all-gathered `[world * max_rows, ld]` table."""
        own = self.owner(ids)
        lo = self._dev(ids.device)[1][own]
        return own * self.max_rows + (ids - lo)

    def unpad(self, table: Tensor) -> Tensor:
        """`[world * max_rows, ...]` -> `[num_nodes, ...]` (drops the padding rows)."""
        parts = [table[r * self.max_rows: r * self.max_rows + self.hi(r) - self.lo(r)] for r in range(self.world)]
        return torch.cat(parts, 0)


# --------------------------------------------------------------------------------- peer memory
class PeerArena:
    """One device arena per rank (allocated and exported by liblgc_b200.so as a CUDA IPC handle), mapped
    into every other rank of `group`: tables carved out of it sit at the same offset on every rank, so
    `lgc_item_exchange` can load the peers' partial sums and store result rows into every replica over
    NVLink. The 64-byte handles travel through one all-gather of `torch.distributed`."""
    ALIGN = 256

    def __init__(self, capi, lib, nbytes: int, device, group=None):
        """Allocates this rank's arena only (local, cannot block); `connect()` maps the peers."""
        self._capi, self._lib, self.device, self.group = capi, lib, torch.device(device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise RuntimeError("peer arenas support up to 8 ranks")
        self.nbytes = (int(nbytes) + self.ALIGN - 1) // self.ALIGN * self.ALIGN + 256      # + control block
        self.ctrl_off = self.nbytes - 256
        self._top = 0
        base, self._handle = C.c_void_p(), (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            capi.check(lib.lgc_peer_arena_alloc(self.nbytes, C.byref(base), self._handle), "lgc_peer_arena_alloc")
        self.base = int(base.value)
        self.bases, self._opened = [0] * self.world, []
        self.bases[self.rank] = self.base

    def connect(self) -> None:
        """Collective: all-gathers the 64-byte IPC handles and maps every peer's arena. Every rank of the
        group must call it (the caller makes sure every rank's allocation succeeded first)."""
        if self.world == 1:
            return
        mine = torch.tensor(list(bytes(self._handle)), dtype=torch.uint8, device=self.device)
        every = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.cpu().numpy().tobytes()
        with torch.cuda.device(self.device):
            for q in range(self.world):
                if q == self.rank:
                    continue
                peer = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(every[q * 64:(q + 1) * 64])
                self._capi.check(self._lib.lgc_peer_arena_open(buf, C.byref(peer)), "lgc_peer_arena_open")
                self.bases[q] = int(peer.value)
                self._opened.append(int(peer.value))

    def table(self, rows: int, ld: int) -> Tensor:
        """A zero-filled fp32 `[rows, ld]` table inside the arena (same offset on every rank when every
        rank asks for the same sequence of tables)."""
        from .graph import _DeviceArray
        n = max(int(rows), 1) * int(ld)
        off = self._top
        self._top = (off + 4 * n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if self._top > self.ctrl_off:
            raise RuntimeError("peer arena exhausted")
        t = torch.as_tensor(_DeviceArray(self.base + off, n, "<f4"), device=self.device).view(max(int(rows), 1), int(ld))
        t._lgc_arena = self                                    # the view must not outlive the arena
        return t

    @staticmethod
    def bytes_for(tables: int, rows: int, ld: int) -> int:
        per = (4 * max(int(rows), 1) * int(ld) + PeerArena.ALIGN - 1) // PeerArena.ALIGN * PeerArena.ALIGN
        return tables * per

    def descriptor(self, part: Tensor, ld: int, timeout_ms: int = 0):
        x = self._capi.PeerExchange(world=self.world, rank=self.rank, arena_bytes=self.nbytes, ctrl_off=self.ctrl_off,
                                    part=part.data_ptr(), n_rows=part.size(0), ld=ld, timeout_ms=timeout_ms)
        for q, b in enumerate(self.bases):
            x.bases[q] = b
        return x

    def status(self) -> Tuple[int, int]:
        """(error word, exchanges completed) of this rank's control block; synchronises the device."""
        x = self._capi.PeerExchange(world=self.world, rank=self.rank, arena_bytes=self.nbytes, ctrl_off=self.ctrl_off)
        for q, b in enumerate(self.bases):
            x.bases[q] = b
        err, epoch = C.c_int32(), C.c_int64()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            self._capi.check(self._lib.lgc_peer_exchange_status(C.byref(x), C.byref(err), C.byref(epoch)),
                             "lgc_peer_exchange_status")
        return int(err.value), int(epoch.value)

    def close(self) -> None:
        """Unmap the peers' arenas and free the own one. Collective: every rank must have finished using
        every arena (the caller synchronises and runs a barrier first)."""
        if self.base == 0:
            return
        with torch.cuda.device(self.device):
            for p in self._opened:
                self._lib.lgc_peer_arena_close(p)
            self._lib.lgc_peer_arena_free(self.base)
        self._opened, self.base = [], 0


# --------------------------------------------------------------------------------- CUDA backend
class CudaBackend:
    """The product backend: raw pointers into liblgc_b200.so. No CPU path."""
    supports_graph = True

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

    def _epilogue(self, mode, y, acc, xrow, addend, a0, a1, scale, beta, p, m, v, lr, betas, eps, step,
                  adam_scalars, hist, ah):
        from .graph import _ptr
        e = self._capi.SpmmEpilogue(mode=mode, a0=a0, a1=a1, scale=scale, beta=beta, y=_ptr(y), acc=_ptr(acc),
                                    xrow=_ptr(xrow), addend=_ptr(addend), p=_ptr(p), m=_ptr(m), v=_ptr(v),
                                    lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, step=step,
                                    adam_scalars=_ptr(adam_scalars))
        if hist is not None:
            e.n_hist = len(hist)
            for i, (t, w) in enumerate(zip(hist, ah)):
                e.hist[i] = _ptr(t)
                e.ah[i] = float(w)
        return e

    def spmm_ex(self, handle, ld: int, x: Tensor, ws: Tensor, mode: int, *, y=None, acc=None, xrow=None,
                addend=None, a0=0.0, a1=0.0, scale=1.0, beta=0.0, p=None, m=None, v=None, lr=0.0,
                betas=(0.9, 0.999), eps=1e-8, step=1, adam_scalars=None, hist=None, ah=None) -> None:
        """One LGConv layer with a fused epilogue (mode: 0 PLAIN, 1 FWD_INIT, 2 FWD_RMW, 3 ADAM, 4 FWD_FINAL)."""
        from .graph import _ptr, _stream
        e = self._epilogue(mode, y, acc, xrow, addend, a0, a1, scale, beta, p, m, v, lr, betas, eps, step,
                           adam_scalars, hist, ah)
        with torch.cuda.device(x.device):
            rc = self._lib.lgc_spmm_ex(handle, ld, _ptr(x), C.byref(e), _ptr(ws), ws.numel(), _stream())
        self._capi.check(rc, "lgc_spmm_ex")

    def epilogue_apply(self, sums: Tensor, ld: int, mode: int, *, y=None, acc=None, xrow=None, addend=None, a0=0.0,
                       a1=0.0, scale=1.0, beta=0.0, p=None, m=None, v=None, lr=0.0, betas=(0.9, 0.999), eps=1e-8,
                       step=1, adam_scalars=None, hist=None, ah=None) -> None:
        """The same epilogue on row sums that already exist (the all-reduced item rows)."""
        from .graph import _ptr, _stream
        e = self._epilogue(mode, y, acc, xrow, addend, a0, a1, scale, beta, p, m, v, lr, betas, eps, step,
                           adam_scalars, hist, ah)
        with torch.cuda.device(sums.device):
            rc = self._lib.lgc_epilogue_apply(sums.size(0), ld, _ptr(sums), C.byref(e), _stream())
        self._capi.check(rc, "lgc_epilogue_apply")

    # ---- peer-memory item exchange (csrc/exchange.cu)
    supports_peer = True

    def peer_arena(self, nbytes: int, device, group) -> "PeerArena":
        return PeerArena(self._capi, self._lib, nbytes, device, group)

    def item_exchange(self, arena: "PeerArena", part: Tensor, ld: int, mode: int, *, y=None, acc=None, addend=None,
                      a1=0.0, scale=1.0, beta=0.0, p=None, m=None, v=None, lr=0.0, betas=(0.9, 0.999), eps=1e-8,
                      step=1, adam_scalars=None, hist=None, ah=None) -> None:
        """Sum of the ranks' partial item sums + fused epilogue + broadcast of the result rows, one launch
        per rank over NVLink peer memory (`lgc_item_exchange`); every rank issues the same sequence."""
        from .graph import _stream
        e = self._epilogue(mode, y, acc, None, addend, 0.0, a1, scale, beta, p, m, v, lr, betas, eps, step,
                           adam_scalars, hist, ah)
        x = arena.descriptor(part, ld)
        with torch.cuda.device(part.device):
            rc = self._lib.lgc_item_exchange(C.byref(x), C.byref(e), _stream())
        self._capi.check(rc, "lgc_item_exchange")

    def row_degree(self, handle, n_rows: int, device) -> Tensor:
        """fp32 weighted in-degree of every row of a rect graph, summed in edge-list order (bit-exact
        to the CPU scatter_add_ of gcn_norm for the edges the graph holds)."""
        from .graph import _from_device_ptr
        info = self._capi.GraphInfo()
        self._capi.check(self._lib.lgc_graph_get_info(handle, C.byref(info)), "lgc_graph_get_info")
        return _from_device_ptr(info.deg, int(info.num_nodes), torch.float32, device)[:n_rows].clone()

    def scatter_add(self, table: Tensor, idx: Tensor, rows: Tensor) -> None:
        """table[idx[j]] += rows[j], duplicates in input order (deterministic); idx < 0 skipped."""
        from .graph import _ptr, _stream
        with torch.cuda.device(table.device):
            rc = self._lib.lgc_scatter_add_rows(idx.numel(), table.stride(0), _ptr(idx.contiguous()),
                                                _ptr(rows.contiguous()), _ptr(table), _stream())
        self._capi.check(rc, "lgc_scatter_add_rows")

    def adam_step(self, p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, betas, eps: float, step: int,
                  adam_scalars=None):
        from . import ops
        if adam_scalars is None:
            ops.adam_step(p, g.contiguous(), m, v, lr, betas, eps, step)
            return
        from .graph import _ptr, _stream
        g = g.contiguous()
        with torch.cuda.device(p.device):
            rc = self._lib.lgc_adam_step_dev(p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(adam_scalars),
                                             _stream())
        self._capi.check(rc, "lgc_adam_step_dev")

    def adam_scalars(self, lr: float, betas, eps: float, step: int) -> Tensor:
        """The 6 step-dependent Adam scalars as a fresh host tensor (float32[6])."""
        arr = (C.c_float * 6)()
        self._capi.check(self._lib.lgc_adam_scalars(lr, betas[0], betas[1], eps, step, arr), "lgc_adam_scalars")
        return torch.tensor(list(arr), dtype=torch.float32)

    def bpr(self, outc: Tensor, e0c: Tensor, batch: int, decay: float, alpha0: float):
        """BPR + L2 on the compact `[3*batch, ld]` row tables (users | pos | neg)."""
        from . import ops
        dev = outc.device
        ar = torch.arange(batch, device=dev, dtype=torch.int64)
        gc, zc = torch.zeros_like(outc), torch.zeros_like(outc)
        loss3 = ops.bpr_loss_grad(outc, e0c, ar, ar + batch, ar + 2 * batch, decay, alpha0, gc, zc)
        return loss3, gc, zc


# --------------------------------------------------------------------------------- CUDA-graph replay
class _GraphedStep:
    """`step()` for both sharded trainers. A sharded step is ~100 small launches (kernels through
    the C ABI, NCCL collectives, index plumbing) issued from Python: beyond 4 GPUs the host cannot
    issue them as fast as the GPUs retire them. So after two eager steps the whole step -- NCCL
    included -- is captured into one CUDA graph and replayed; the only per-step host work is copying
    the triples into static buffers and refreshing the six Adam scalars (launch arguments are frozen
    in a graph, so the update kernels read them from device memory)."""
    use_graph = True

    def _adam_kw(self):
        return {"adam_scalars": self._adam_dev} if self._adam_dev is not None else {}

    def _init_graph_state(self):
        self._graph, self._graph_key, self._warm = None, None, 0
        self._graphable = bool(getattr(self.backend, "supports_graph", False)) and self.dev.type == "cuda"
        self._adam_dev = None
        if self._graphable:
            self._adam_dev = torch.zeros(6, dtype=torch.float32, device=self.dev)

    def step(self, users: Tensor, pos: Tensor, neg: Tensor, decay: float) -> Tensor:
        self.step_count += 1
        users, pos, neg = (t.to(device=self.dev, dtype=torch.int64) for t in (users, pos, neg))
        if not self._graphable:
            return self._step_impl(users, pos, neg, decay)
        # a fresh pageable host tensor per step: the copy is staged before this call returns, so the
        # host running ahead of the GPU cannot overwrite scalars a queued step has not read yet
        self._adam_dev.copy_(self.backend.adam_scalars(self.lr, self.betas, self.eps, self.step_count))
        key = (users.numel(), float(decay))
        if not self.use_graph or (self._graph_key != key and self._warm < 2):
            self._warm += 1
            return self._step_impl(users, pos, neg, decay)
        if self._graph_key != key:
            self._static = [users.clone(), pos.clone(), neg.clone()]
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._static_loss = self._step_impl(*self._static, decay)
            self._graph, self._graph_key = graph, key
        else:
            for dst, src in zip(self._static, (users, pos, neg)):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static_loss.clone()

    def release_graph(self) -> None:
        """Drop the captured graph (call before destroying the process group)."""
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        self._graph, self._graph_key, self._static_loss = None, None, None


# --------------------------------------------------------------------------------- trainer
class ShardedBPRTrainer(_GraphedStep):
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
        self._init_graph_state()

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
    def _step_impl(self, users: Tensor, pos: Tensor, neg: Tensor, decay: float) -> Tensor:
        b, a, K, ld = self.backend, self.alpha, self.layers, self.ld
        batch = users.numel()
        self.propagate()

        # ---- the <= 3*batch needed rows of out / E0: owners contribute, one small all-reduce
        # (fixed shapes, no host synchronisation: rows of other ranks are masked to zero)
        ids = torch.cat([users, pos, neg]).to(device=self.dev, dtype=torch.int64)
        mine = (ids >= self.lo) & (ids < self.hi)
        loc = torch.where(mine, ids - self.lo, -1)
        locc = loc.clamp_min(0)
        rows = torch.stack([self.out[locc], self.e0[locc]]) * mine[None, :, None]
        self._all_reduce(rows)
        loss3, gc, zc = b.bpr(rows[0], rows[1], batch, float(decay), a[0])

        # ---- gradient rows: dL/d out for everybody (gather source of the backward), Z for owners
        pid = self.part.padded_id(ids)
        b.scatter_add(self.gfull, pid, gc)
        b.scatter_add(self.z, loc, zc)
        g_local = self.gfull[self.rank * self.max_rows: (self.rank + 1) * self.max_rows]

        # ---- backward (Horner on the symmetric operator) + Adam on the owned rows
        cur, scale = self.gfull, a[K]
        for l in range(K - 1, 0, -1):                         # h_l = alpha_l G + A h_{l+1}
            b.spmm_ex(self.handle, ld, cur, self.ws, 0, y=self.y, addend=g_local, scale=scale, beta=a[l])
            cur = self.full[l & 1]
            self._all_gather(cur, self.y)
            scale = 1.0
        b.spmm_ex(self.handle, ld, cur, self.ws, 3, addend=self.z, scale=scale, p=self.e0, m=self.m, v=self.v,
                  lr=self.lr, betas=self.betas, eps=self.eps, step=self.step_count, **self._adam_kw())
        self.gfull.index_fill_(0, pid, 0.0)
        self.z.index_fill_(0, locc, 0.0)
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


# --------------------------------------------------------------------------------- bipartite-aware
def bipartite_split(edge_index: Tensor) -> Optional[int]:
    """`s` such that every edge joins a node < s (a user) with a node >= s (an item) -- the layout
    `df_to_graph` produces (items offset by n_users, reference `src/utils_v2.py:128`) -- else None."""
    if edge_index.numel() == 0:
        return None
    s = int(torch.minimum(edge_index.max(dim=0).values.min(), edge_index.max()).item())
    lo_side = edge_index < s
    return s if bool((lo_side[0] != lo_side[1]).all()) and s > 0 else None


class _PendingItems:
    """The item rows of one layer, in flight (see `BipartiteShardedTrainer._item_partials`)."""

    def __init__(self, finish):
        self._finish = finish

    def wait(self) -> None:
        f, self._finish = self._finish, None
        if f is not None:
            f()


class BipartiteShardedTrainer(_GraphedStep):
    """Bipartite-aware sharding of the same step (~30x less traffic than the
    all-gather of whole tables): USERS are partitioned over the ranks (balanced by in-degree + 4),
    the small ITEM table is replicated. Per layer a rank computes its users' rows from the replicated
    item table (rows kernel, fused epilogue, no communication) and the PARTIAL sums of every item row
    over its own users (sweep kernel). The item rows are completed by `lgc_item_exchange`
    (`csrc/exchange.cu`, `exchange="peer"`): ONE kernel per rank and layer over NVLink peer memory that
    sums a row slice over all ranks' partial tables (P2P loads, fixed rank order), applies the item
    rows' fused epilogue and stores the results into every rank's replica (P2P stores) -- or, with
    `exchange="nccl"`, by an asynchronous `ncclAllReduce` of the `[n_items, ld]` partials followed by
    one launch of the same epilogue (`lgc_epilogue_apply`) on every rank. Either way the exchange of
    layer l runs on a side stream while the rank computes its user rows of layer l and the item
    partials of layer l + 1. Item-row sums are reduced in a different order than on one GPU: equal
    within fp32 tolerance, not bit-exact (the two exchange paths agree bit for bit at two ranks).

    A rank needs only ITS OWN users' interactions: the two rectangular operators are built from that
    slice (`from_pairs`), the item degrees are the all-reduced partial degrees, so no process ever
    holds the global edge list (c5: 200 M interactions)."""

    def __init__(self, edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int, embedding_dim: int,
                 num_layers: int, init_weight: Tensor, n_users: int, lr: float = 0.005, betas=(0.9, 0.999),
                 eps: float = 1e-8, alpha: Optional[Sequence[float]] = None, group=None, backend=None,
                 ld: Optional[int] = None, exchange: str = "auto"):
        """From the reference's global edge list (`df_to_graph` layout: [[u; i], [i; u]], every rank passes
        the same one): the first half holds every interaction once, in frame order."""
        n_users = int(n_users)
        assert int(edge_index[1].min().item()) >= 0
        if not self._pairs_layout_ok(edge_index, n_users):
            raise RuntimeError("the sharded step needs a symmetric graph (backward reuses the operator)")
        e = edge_index.size(1) // 2
        user, item = edge_index[0, :e], edge_index[1, :e] - n_users
        w = edge_weight[:e] if edge_weight is not None else torch.ones(e, dtype=torch.float32, device=edge_index.device)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        deg_cnt = torch.bincount(user, minlength=n_users)
        part = RowPartition(deg_cnt.cpu().numpy(), world)
        lo, hi = part.lo(rank), part.hi(rank)
        mine = (user >= lo) & (user < hi)                                    # keeps frame order
        self._setup(user[mine], item[mine], w[mine].float(), part, n_users, int(num_nodes) - n_users, embedding_dim,
                    num_layers, init_weight[lo:hi], init_weight[n_users:], lr, betas, eps, alpha, group, backend, ld,
                    exchange)

    @staticmethod
    def _pairs_layout_ok(edge_index: Tensor, n_users: int) -> bool:
        if edge_index.size(1) % 2:
            return False
        e = edge_index.size(1) // 2
        return bool(torch.equal(edge_index[0, :e], edge_index[1, e:]) and torch.equal(edge_index[1, :e], edge_index[0, e:])
                    and (edge_index[0, :e] < n_users).all() and (edge_index[1, :e] >= n_users).all())

    @classmethod
    def from_pairs(cls, user: Tensor, item: Tensor, weight: Tensor, part: RowPartition, n_users: int, n_items: int,
                   embedding_dim: int, num_layers: int, init_users: Tensor, init_items: Tensor, lr: float = 0.005,
                   betas=(0.9, 0.999), eps: float = 1e-8, alpha=None, group=None, backend=None, ld=None,
                   exchange: str = "auto"):
        """From this rank's OWN interactions only: `user` (global user ids inside the rank's range of
        `part`), `item` (un-offset item ids), `weight`; `init_users` = the rank's rows of the initial
        table, `init_items` = all item rows (replicated)."""
        self = cls.__new__(cls)
        self._setup(user, item, weight.float(), part, n_users, n_items, embedding_dim, num_layers, init_users,
                    init_items, lr, betas, eps, alpha, group, backend, ld, exchange)
        return self

    def _setup(self, user, item, w, part, n_users, n_items, embedding_dim, num_layers, init_users, init_items, lr,
               betas, eps, alpha, group, backend, ld, exchange="auto"):
        assert num_layers >= 1
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = backend if backend is not None else CudaBackend()
        self.dev = user.device
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.num_nodes, self.dim, self.layers = self.n_users + self.n_items, int(embedding_dim), int(num_layers)
        if ld is None:
            from .graph import padded_dim
            ld = padded_dim(self.dim)
        self.ld = ld
        self.alpha = [float(a) for a in (alpha if alpha is not None else [1.0 / (num_layers + 1)] * (num_layers + 1))]
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self.part = part
        lo, hi = part.lo(self.rank), part.hi(self.rank)
        self.lo, self.hi, self.n_local, self.max_rows = lo, hi, hi - lo, part.max_rows
        b, nu, ni = self.backend, self.max_rows, self.n_items
        ul = user - lo

        # ---- gcn_norm without the global graph: weighted in-degrees from the two local operators built on
        # the RAW weights (users: all of a user's edges are local -> bit-exact; items: partial sums,
        # all-reduced), then w_hat = (dis[src] * w) * dis[dst] per interaction, the same for both directions
        g_u = b.build_rect(item, ul, w, max(self.n_local, 1), ni)
        g_i = b.build_rect(ul, item, w, ni, nu)
        deg_u = b.row_degree(g_u, max(self.n_local, 1), self.dev)
        deg_i = b.row_degree(g_i, ni, self.dev)
        b.destroy(g_u); b.destroy(g_i)
        if self.world > 1:
            dist.all_reduce(deg_i, group=self.group)
        dis_u, dis_i = deg_u.pow(-0.5), deg_i.pow(-0.5)
        dis_u[torch.isinf(dis_u)] = 0
        dis_i[torch.isinf(dis_i)] = 0
        self.dis_u, self.dis_i = dis_u, dis_i                  # deg^-1/2 (fp32, like gcn_norm): kept for checks
        w_ui = (dis_u[ul] * w) * dis_i[item]                   # edge user -> item (target item)
        w_iu = (dis_i[item] * w) * dis_u[ul]                   # edge item -> user (target user)
        self.gu = b.build_rect(item, ul, w_iu, max(self.n_local, 1), ni)
        self.gi = b.build_rect(ul, item, w_ui, ni, nu)         # user tables have max_rows rows
        self.local_nnz = 2 * int(user.numel())
        del w_ui, w_iu, ul
        self.ws_u = b.workspace(self.gu, ld, self.dev)
        self.ws_i = b.workspace(self.gi, ld, self.dev)

        def table(rows):
            return torch.zeros(max(rows, 1), ld, dtype=torch.float32, device=self.dev)
        n_x = max(self.layers - 1, 2)                           # stored layers (forward), ping-pong (backward)
        # ---- item-row exchange: one kernel over NVLink peer memory (`lgc_item_exchange`) when the backend
        # has it, else ncclAllReduce + `lgc_epilogue_apply`. Every item table an epilogue writes and the
        # partial-sum tables then live in the rank's peer arena, at the same offsets on every rank.
        self.peer = self._open_peer_arena(exchange, PeerArena.bytes_for(6 + n_x, ni, ld))
        itable = self.peer.table if self.peer is not None else (lambda rows, ld_: table(rows))
        self.e0_u, self.m_u, self.v_u = table(nu), table(nu), table(nu)
        self.e0_i, self.m_i, self.v_i = itable(ni, ld), itable(ni, ld), itable(ni, ld)
        self.e0_u[: self.n_local, : self.dim] = init_users.to(self.dev)
        self.e0_i[:, : self.dim] = init_items.to(self.dev)
        self.out_u, self.out_i = table(nu), itable(ni, ld)
        self.xu, self.xi = [table(nu) for _ in range(n_x)], [itable(ni, ld) for _ in range(n_x)]
        self.part_i = [itable(ni, ld), itable(ni, ld)]          # partial item sums in flight (two layers)
        self.g_u, self.z_u, self.g_i, self.z_i = table(nu), table(nu), table(ni), table(ni)
        self.n_cols = self.n_items                              # rows exchanged per layer
        self._side = torch.cuda.Stream(self.dev) if self.peer is not None else None
        if self.peer is not None:
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.group)                      # every arena is initialised before any peer writes
        self._init_graph_state()

    def _open_peer_arena(self, exchange: str, nbytes: int):
        """"peer": required; "nccl": never; "auto": when the backend has it, 2..8 ranks, and every rank
        could map every arena (else all ranks fall back to NCCL together). LGC_EXCHANGE overrides "auto"."""
        import os
        if exchange == "auto":
            exchange = os.environ.get("LGC_EXCHANGE", "auto")
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        # (`peer_on_cpu`: only the CPU test double sets it, to drive the collective agreement below under gloo)
        able = (exchange != "nccl" and self.world > 1 and self.world <= 8
                and (self.dev.type == "cuda" or bool(getattr(self.backend, "peer_on_cpu", False)))
                and bool(getattr(self.backend, "supports_peer", False)))
        if not able:
            if exchange == "peer":
                raise RuntimeError("exchange='peer' needs 2..8 CUDA ranks and a backend with lgc_item_exchange")
            return None
        # two collective phases, each agreed on by all ranks before the next starts: a rank that fails must
        # not leave the others waiting in the all-gather of the handles
        def all_ok(flag: bool) -> bool:
            ok = torch.tensor([1 if flag else 0], device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            return int(ok.item()) == 1

        arena, err = None, None
        try:
            arena = self.backend.peer_arena(nbytes, self.dev, self.group)        # local allocation + IPC export
        except Exception as e:                                   # noqa: BLE001 - reported below, collectively
            err = e
        if all_ok(arena is not None):
            try:
                arena.connect()                                  # all-gather of the handles + peer mappings
            except Exception as e:                               # noqa: BLE001
                err = e
            if all_ok(err is None):
                return arena
        if arena is not None:
            arena.close()
        if exchange == "peer":
            raise RuntimeError(f"peer arenas could not be mapped on every rank: {err}")
        import warnings
        warnings.warn(f"lgc_item_exchange unavailable ({err}); item rows go through ncclAllReduce")
        return None

    def check_exchange(self) -> None:
        """Raises if a barrier of the peer-memory exchange timed out (synchronises the device)."""
        if self.peer is not None:
            err, _ = self.peer.status()
            if err:
                raise RuntimeError(f"lgc_item_exchange: a peer did not arrive in time (phase {err}); results invalid")

    def close(self) -> None:
        """Collective teardown of the peer arenas (call on every rank before destroying the process group).
        The item tables live inside the arena: they are moved to ordinary tensors first, so `weight()` and
        `embedding()` (through the NCCL path) stay valid afterwards; `step()` does not -- with the peer
        exchange the Adam moments of the item rows are sharded by row slice, not replicated."""
        self.release_graph()
        if getattr(self, "peer", None) is not None:
            if self.dev.type == "cuda":
                torch.cuda.synchronize(self.dev)
            if dist.is_initialized():
                dist.barrier(group=self.group)               # nobody reads or writes any arena after this point
            for name in ("e0_i", "m_i", "v_i", "out_i"):
                setattr(self, name, getattr(self, name).clone())
            self.xi = [t.clone() for t in self.xi]
            self.part_i = [t.clone() for t in self.part_i]
            if self.dev.type == "cuda":
                torch.cuda.synchronize(self.dev)
            self.peer.close()
            self.peer, self._side, self._closed = None, None, True

    def __del__(self):
        for name in ("gu", "gi"):
            h = getattr(self, name, None)
            setattr(self, name, None)
            if h is not None:
                try:
                    self.backend.destroy(h)
                except Exception:
                    pass

    def _all_reduce(self, t: Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, group=self.group)

    def _item_partials(self, x_u: Tensor, part: Tensor, mode: int, **epi) -> "_PendingItems":
        """part = this rank's partial sums of (A_hat x)[items] over its own users, then -- ASYNCHRONOUSLY --
        their sum over the ranks and the item rows' epilogue `mode` (0 PLAIN into epi['y'], 3 ADAM, 4
        FWD_FINAL). The caller waits on the returned object only when it needs the item rows, so the NVLink
        exchange hides behind the kernels issued in between. Peer mode: one `lgc_item_exchange` launch on a
        side stream does all of it; NCCL mode: async all-reduce, epilogue on the main stream at wait()."""
        b = self.backend
        b.spmm_ex(self.gi, self.ld, x_u, self.ws_i, 0, y=part, scale=1.0)
        if self.peer is not None:
            main = torch.cuda.current_stream(self.dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                b.item_exchange(self.peer, part, self.ld, mode, **epi)
            return _PendingItems(lambda: main.wait_stream(self._side))
        work = dist.all_reduce(part, group=self.group, async_op=True) if self.world > 1 else None
        plain_in_place = mode == 0 and epi.get("y") is part     # forward layers < K: the sums ARE x_l

        def finish():
            if work is not None:
                work.wait()
            if not plain_in_place:
                b.epilogue_apply(part, self.ld, mode, **epi)
        return _PendingItems(finish)

    # ------------------------------------------------------------------ forward only
    def propagate(self) -> Tuple[Tensor, Tensor]:
        """out = sum_l alpha_l A_hat^l E0: user rows [max_rows, ld] (own users), item rows (replicated).
        Layers 1..K-1 store x_l; layer K folds the whole mean into its epilogue (FWD_FINAL), for the
        user rows inside the SpMM, for the item rows in one `lgc_epilogue_apply` after the all-reduce."""
        b, a, K, ld = self.backend, self.alpha, self.layers, self.ld
        hist_u = [self.e0_u]
        hist_i = [self.e0_i] + self.xi[: K - 1]                 # layer tables of the items, in layer order

        def items_of_layer(l: int, x_u: Tensor) -> "_PendingItems":
            if l < K:                                           # x_l of the items, summed in place
                return self._item_partials(x_u, self.xi[l - 1], 0, y=self.xi[l - 1], scale=1.0)
            return self._item_partials(x_u, self.part_i[0], 4, acc=self.out_i, a1=a[K], hist=hist_i, ah=a[:K])

        ci = self.e0_i
        pend = items_of_layer(1, self.e0_u)
        for l in range(1, K + 1):
            last = l == K
            # user rows of layer l from the item rows of layer l-1
            if last:
                b.spmm_ex(self.gu, ld, ci, self.ws_u, 4, acc=self.out_u, a1=a[K], hist=hist_u, ah=a[:K])
            else:
                b.spmm_ex(self.gu, ld, ci, self.ws_u, 0, y=self.xu[l - 1], scale=1.0)
            # item partials of layer l+1 need only the user rows of layer l: issued BEFORE waiting for the
            # exchange of layer l, which therefore overlaps both kernels
            npend = None if last else items_of_layer(l + 1, self.xu[l - 1])
            pend.wait()
            if not last:
                ci = self.xi[l - 1]                             # x_l of the items
                hist_u.append(self.xu[l - 1])
                pend = npend
        return self.out_u, self.out_i

    # ------------------------------------------------------------------ one mini-batch
    def _step_impl(self, users: Tensor, pos: Tensor, neg: Tensor, decay: float) -> Tensor:
        if getattr(self, "_closed", False):
            raise RuntimeError("the trainer was closed (its peer arenas are gone): build a new one to keep training")
        b, a, K, ld, s = self.backend, self.alpha, self.layers, self.ld, self.n_users
        batch = users.numel()
        self.propagate()

        users = users.to(device=self.dev, dtype=torch.int64)
        items = torch.cat([pos, neg]).to(device=self.dev, dtype=torch.int64) - s
        mine = (users >= self.lo) & (users < self.hi)
        loc = torch.where(mine, users - self.lo, -1)            # fixed shapes, no host synchronisation
        locc = loc.clamp_min(0)
        urows = torch.stack([self.out_u[locc], self.e0_u[locc]]) * mine[None, :, None]
        self._all_reduce(urows)                              # the batch's user rows; item rows are local
        outc = torch.cat([urows[0], self.out_i[items]])
        e0c = torch.cat([urows[1], self.e0_i[items]])
        loss3, gc, zc = b.bpr(outc, e0c, batch, float(decay), a[0])
        b.scatter_add(self.g_u, loc, gc[:batch])
        b.scatter_add(self.z_u, loc, zc[:batch])
        b.scatter_add(self.g_i, items, gc[batch:])           # replicated: bit-identical on every rank
        b.scatter_add(self.z_i, items, zc[batch:])

        # ---- backward (Horner on the symmetric operator) + Adam: h_l = alpha_l G + A h_{l+1}
        adam = dict(lr=self.lr, betas=self.betas, eps=self.eps, step=self.step_count, **self._adam_kw())
        def items_bwd(l: int, x_u: Tensor, scale: float) -> "_PendingItems":
            part = self.part_i[(K - 1 - l) & 1]
            if l > 0:
                return self._item_partials(x_u, part, 0, y=self.xi[l & 1], addend=self.g_i, scale=scale, beta=a[l])
            return self._item_partials(x_u, part, 3, addend=self.z_i, scale=scale, p=self.e0_i, m=self.m_i,
                                       v=self.v_i, **adam)

        ci, scale = self.g_i, a[K]
        pend = items_bwd(K - 1, self.g_u, scale)
        for l in range(K - 1, -1, -1):
            if l > 0:
                nu_ = self.xu[l & 1]
                b.spmm_ex(self.gu, ld, ci, self.ws_u, 0, y=nu_, addend=self.g_u, scale=scale, beta=a[l])
                npend = items_bwd(l - 1, nu_, 1.0)              # item partials of the next layer
                pend.wait()
                ci, scale, pend = self.xi[l & 1], 1.0, npend
            else:
                b.spmm_ex(self.gu, ld, ci, self.ws_u, 3, addend=self.z_u, scale=scale, p=self.e0_u, m=self.m_u,
                          v=self.v_u, **adam)
                pend.wait()
        self.g_u.index_fill_(0, locc, 0.0)
        self.z_u.index_fill_(0, locc, 0.0)
        self.g_i.index_fill_(0, items, 0.0)
        self.z_i.index_fill_(0, items, 0.0)
        return loss3

    # ------------------------------------------------------------------ views for callers / tests
    def _gather_users(self, shard: Tensor) -> Tensor:
        full = torch.empty(self.world * self.max_rows, self.ld, dtype=torch.float32, device=self.dev)
        if self.world == 1:
            full.copy_(shard)
        else:
            dist.all_gather_into_tensor(full, shard, group=self.group)
        return self.part.unpad(full)

    def weight(self) -> Tensor:
        self.check_exchange()
        return torch.cat([self._gather_users(self.e0_u), self.e0_i])[:, : self.dim]

    def embedding(self) -> Tensor:
        self.propagate()
        self.check_exchange()
        return torch.cat([self._gather_users(self.out_u), self.out_i])[:, : self.dim]


def make_sharded_trainer(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int, embedding_dim: int,
                         num_layers: int, init_weight: Tensor, mode: str = "auto", **kw):
    """`mode`: "bipartite" (users sharded, items replicated + all-reduced), "rows" (destination rows
    partitioned, all-gather per layer -- works for any symmetric graph) or "auto" (bipartite when the
    edge list has the users-then-items layout). `exchange` ("auto" | "peer" | "nccl", bipartite only): how the
    item rows are completed per layer -- see `BipartiteShardedTrainer`."""
    s = bipartite_split(edge_index) if mode in ("auto", "bipartite") else None
    if mode == "bipartite" and s is None:
        raise ValueError("edge_index is not bipartite with users numbered before items")
    if s is not None:
        return BipartiteShardedTrainer(edge_index, edge_weight, num_nodes, embedding_dim, num_layers, init_weight,
                                       n_users=s, **kw)
    kw.pop("exchange", None)              # only the bipartite trainer has item rows to exchange
    return ShardedBPRTrainer(edge_index, edge_weight, num_nodes, embedding_dim, num_layers, init_weight, **kw)
