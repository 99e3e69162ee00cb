"""Runs the reference's OWN `networks.py` / `data.py` (verbatim copies in the git-ignored `oracle/_ref/`, made by
`tools/make_oracle_ref.sh`) over `oracle/pyg_shim` - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used by `bench.py --impl reference` / its `cpu_baseline` leg (kind "reference") and by tests that execute the reference
script.  The product package never imports this file.  `available()` is False on a machine where the recipe was never run
(then bench.py falls back to the oracle port and says so).

Only one thing of the reference is neutralised, and only when asked (`stub_laplacian=True`): `data.get_laplacian`, whose
result the loss discards (data.py:422-423) but whose dense form is O(Nt^2) - 329 GB at Oberrhein B=4096 (SURVEY.md headline 3).
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SHIM_DIR = os.path.join(HERE, "pyg_shim")
_cache = {}


def available():
    return os.path.exists(os.path.join(REF_DIR, "networks.py")) and os.path.exists(os.path.join(REF_DIR, "data.py"))


def load_reference(stub_laplacian=False):
    """(ref_networks, ref_data): the reference modules under private names (they do not shadow the product's modules)."""
    if not available():
        raise RuntimeError("oracle/_ref is empty: run tools/make_oracle_ref.sh where the reference checkout exists")
    if "mods" not in _cache:
        if SHIM_DIR not in sys.path:
            sys.path.insert(0, SHIM_DIR)
        mods = []
        for name in ("networks", "data"):
            spec = importlib.util.spec_from_file_location(f"_dss2_reference_{name}", os.path.join(REF_DIR, f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods.append(mod)
        _cache["mods"] = tuple(mods)
    net, dat = _cache["mods"]
    if stub_laplacian:
        dat.get_laplacian = lambda edge_index=None, **kw: (torch.zeros(2, 1, dtype=torch.long), torch.zeros(1))
    return net, dat


class ReferenceTrainer:
    """dss2_run.py:85-92,131-147 with the SkipPFN line (dss2_run.py:88) selected: model, Adamax(lr 3e-3), one step per call."""

    def __init__(self, state_dict=None, reg_coefs=None, lr=3e-3, stub_laplacian=True,
                 ctor=(8, 6, 2, 32, 8, 2, 0.3, 5)):
        net, dat = load_reference(stub_laplacian=stub_laplacian)
        self.net, self.dat = net, dat
        self.model = net.SkipPFN(*ctor)
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=True)
        self.opt = torch.optim.Adamax(self.model.parameters(), lr=lr)
        self.reg = reg_coefs

    def step(self, batch, stats):
        """batch: dict/obj with x[Nt,11], edge_index, edge_attr[Et,13]; stats: (x_mean, x_std, edge_mean, edge_std)."""
        x, ei, ea = batch["x"], batch["edge_index"], batch["edge_attr"]
        self.model.train()
        self.opt.zero_grad()
        out = self.model(x[:, :8], ei, ea[:, :6])
        loss = self.dat.gsp_wls_edge(input=x[:, :8], edge_input=ea[:, :6], output=out, x_mean=stats[0], x_std=stats[1],
                                     edge_mean=stats[2], edge_std=stats[3], edge_index=ei, reg_coefs=self.reg,
                                     num_samples=None, node_param=x[:, 8:], edge_param=ea[:, 6:])
        loss.backward()
        self.opt.step()
        return loss.detach()
