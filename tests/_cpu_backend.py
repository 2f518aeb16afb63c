"""Test double for `gnn_ecommerce_b200.sharded`: the backend interface of `CudaBackend` restated
with torch CPU ops (the oracle's arithmetic), so the multi-rank orchestration -- partition, id
renumbering, collectives, Horner backward -- can run under `gloo` without a GPU. Test
infrastructure only: nothing in the product imports this module."""
import math

import numpy as np
import torch

from oracle import port


class CpuCheckBackend:
    def global_w_hat(self, edge_index, edge_weight, num_nodes):
        ew = edge_weight if edge_weight is not None else torch.ones(edge_index.size(1))
        csr = port.csr_by_target(edge_index.numpy(), ew.numpy(), num_nodes)
        return torch.from_numpy(csr["w_hat_edge"].copy()), torch.from_numpy(csr["count_deg"].copy()), True

    def build_rect(self, src, dst, w, n_rows, n_cols):
        assert src.numel() == 0 or (int(src.max()) < n_cols and int(dst.max()) < n_rows)
        return {"src": src.clone(), "dst": dst.clone(), "w": w.clone(), "n_rows": n_rows, "n_cols": n_cols}

    def destroy(self, handle):
        pass

    def workspace(self, handle, ld, device):
        return torch.empty(1)

    def row_degree(self, h, n_rows, device):
        return torch.zeros(h["n_rows"]).index_add_(0, h["dst"], h["w"])[:n_rows]    # sequential: edge order

    def spmm_ex(self, h, ld, x, ws, mode, **kw):
        n = h["n_rows"]
        assert x.shape == (h["n_cols"], ld)
        s = torch.zeros(n, ld).index_add_(0, h["dst"], h["w"][:, None] * x[h["src"]])
        self.epilogue_apply(s, ld, mode, **kw)

    def epilogue_apply(self, s, ld, mode, *, y=None, acc=None, xrow=None, addend=None, a0=0.0, a1=0.0, scale=1.0,
                       beta=0.0, p=None, m=None, v=None, lr=0.0, betas=(0.9, 0.999), eps=1e-8, step=1,
                       adam_scalars=None, hist=None, ah=None):
        n = s.size(0)
        f = torch.float32
        if mode == 4:
            r = hist[0][:n] * torch.tensor(ah[0], dtype=f)
            for t, w in zip(hist[1:], ah[1:]):
                r = r + t[:n] * torch.tensor(w, dtype=f)
            acc[:n] = r + s * torch.tensor(a1, dtype=f)
            return
        if mode == 0:
            r = torch.tensor(scale, dtype=f) * s
            if addend is not None:
                r = r + torch.tensor(beta, dtype=f) * addend[:n]
            y[:n] = r
        elif mode == 1:
            if y is not None:
                y[:n] = s
            acc[:n] = xrow[:n] * torch.tensor(a0, dtype=f) + s * torch.tensor(a1, dtype=f)
        elif mode == 2:
            if y is not None:
                y[:n] = s
            acc[:n] = acc[:n] + s * torch.tensor(a1, dtype=f)
        else:
            g = torch.tensor(scale, dtype=f) * s + addend[:n]
            b1, b2 = betas
            bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
            m[:n] = m[:n] + torch.tensor(1 - b1, dtype=f) * (g - m[:n])
            v[:n] = v[:n] * torch.tensor(b2, dtype=f) + torch.tensor(1 - b2, dtype=f) * g * g
            denom = v[:n].sqrt() / torch.tensor(math.sqrt(bc2), dtype=f) + torch.tensor(eps, dtype=f)
            p[:n] = p[:n] + torch.tensor(-(lr / bc1), dtype=f) * m[:n] / denom

    def scatter_add(self, table, idx, rows):
        keep = idx >= 0
        table.index_add_(0, idx[keep], rows[keep])        # CPU index_add_ is sequential: deterministic

    def adam_step(self, p, g, m, v, lr, betas, eps, step, adam_scalars=None):
        f = torch.float32
        b1, b2 = betas
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        m.copy_(m + torch.tensor(1 - b1, dtype=f) * (g - m))
        v.copy_(v * torch.tensor(b2, dtype=f) + torch.tensor(1 - b2, dtype=f) * g * g)
        denom = v.sqrt() / torch.tensor(math.sqrt(bc2), dtype=f) + torch.tensor(eps, dtype=f)
        p.add_(torch.tensor(-(lr / bc1), dtype=f) * m / denom)

    def bpr(self, outc, e0c, batch, decay, alpha0):
        o = outc.clone().requires_grad_(True)
        u, pz, ng = o[:batch], o[batch:2 * batch], o[2 * batch:]
        bpr = torch.nn.functional.softplus(-((u * pz).sum(-1) - (u * ng).sum(-1))).mean()
        bpr.backward()
        reg = decay * 0.5 * e0c.pow(2).sum() / batch
        gc = o.grad
        zc = alpha0 * gc + (decay / batch) * e0c
        return torch.stack([bpr.detach(), reg, bpr.detach() + reg]).float(), gc, zc
