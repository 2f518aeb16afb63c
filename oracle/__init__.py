"""CPU oracle for the LightGCN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product. Only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and there only as the
checker or as the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status: **pinned against the reference's own code, unpinned against real PyG.**
The reference (`/root/reference/src/lightgcn.py`) holds no tests and no golden vectors, and its
propagation arithmetic lives in an absent, un-pinned third-party dependency
(`torch_geometric.nn.conv.LGConv`, `requirements.txt:10`; era-consistent version PyG 2.2.x/2.3.x
for torch==1.13.1). `oracle/lgconv.py` restates that operator's published algorithm op for op
(`lg_conv.py::LGConv.forward` -> `gcn_conv.py::gcn_norm` -> `MessagePassing.propagate` ->
`torch_scatter.scatter_sum`), and `oracle/reference_shim.py` imports the reference's own
`LightGCN`, `BPRLoss` and `utils_v2` UNMODIFIED around it (only possible where
`/root/reference` exists). `tests/golden/make_golden.py` ran that combination in the build
container and committed its outputs as `tests/golden/*.npz`; `oracle/port.py` (the restatement
that travels to the GPU box) is checked against those vectors by the CPU test-suite, and the
`LGConv` restatement is cross-checked against a dense fp64 `D^-1/2 A D^-1/2` product.
"""
