"""Minimal pure-PyTorch stand-in for the handful of `torch_geometric` symbols DSS2 imports.

TEST INFRASTRUCTURE ONLY (part of `oracle/`).  `torch_geometric` is an un-vendored, un-pinned
third-party dependency of the reference (reference README.md:31) that is absent from this image and
from the GPU boxes.  This package restates, from PyG's published semantics, exactly the symbols named
by the reference imports:

    data.py:4,6      Data, InMemoryDataset, download_url, scatter, get_laplacian
    networks.py:4-9  torch_geometric.nn (Sequential), the conv classes, MessagePassing, degree
    dss2_run.py:18   torch_geometric.loader.DataLoader

so that the reference's own `networks.py` / `data.py` can be executed verbatim in this container to
generate the golden vectors under `tests/golden/` (see `tests/golden/make_golden.py`).

Parity status: "parity unpinned" against real PyG - the reference ships no tests or golden vectors
for any PyG-backed piece, and PyG itself cannot be installed here.  Every function
below states the upstream behaviour it encodes so a reviewer with PyG can cross-check.

Nothing in the product path may import this package.
"""
__version__ = "0.0-dss2-oracle-shim"
