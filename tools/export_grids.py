"""One-off exporter (run in the build container, where /root/reference is mounted).

Converts the reference's pandas pickles into plain .npz arrays that can travel to the GPU box
(no pandas objects, no reference code):

  <pkg>/dss2/grids/<case>.npz   bus_param[N,3], edge_param[E_all,9], noise_param[6] (+ column names)
                                from reference data/<case>/{bus_param,edge_param,noise_param}
  tests/golden/cigre14_scenarios.npz  first 128 of the 720 pandapower-solved CIGRE-14 scenarios
                                (nodes[S,15,7], edges[S,17,18], labels[S,15,2], float64) from
                                reference data/cigre14/{nodes,edges,labels}

Column orders are the reference's own (data.py:171-172, toy_network.py:215-229).
"""
import os
import pickle
import sys

import numpy as np

REF = "/root/reference/data"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200", "dss2", "grids")
NOISE_COLS = ["p_noise", "v_noise", "i_noise", "pm_noise", "sgen_noise", "zero_inj_coef"]


def load(case, name):
    with open(os.path.join(REF, case, name), "rb") as fh:
        return pickle.load(fh)


def main():
    for case in ("cigre14", "cigre14_reswitched", "ober_sub"):
        bus, edge, noise = load(case, "bus_param"), load(case, "edge_param"), load(case, "noise_param")
        np.savez_compressed(
            os.path.join(PKG, f"{case}.npz"),
            bus_param=bus[["vn_kv", "bool_slack", "bool_zero_inj"]].values.astype(np.float64),
            edge_param=edge[["from_bus", "to_bus", "G", "B", "Gs", "Bs", "closed line", "phase shift", "imax or sn"]].values.astype(np.float64),
            noise_param=noise[NOISE_COLS].values.astype(np.float64).reshape(-1),
            noise_cols=np.array(NOISE_COLS),
        )
        print(case, bus.shape, edge.shape)
    S = 128
    nodes, edges, labels = load("cigre14", "nodes"), load("cigre14", "edges"), load("cigre14", "labels")
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "cigre14_scenarios.npz"),
        nodes=np.stack([d.values for d in nodes[:S]]).astype(np.float64),
        edges=np.stack([d.values for d in edges[:S]]).astype(np.float64),
        labels=np.stack([d.values for d in labels[:S]]).astype(np.float64),
        node_cols=np.array(list(nodes[0].columns)), edge_cols=np.array(list(edges[0].columns)),
    )
    print("scenarios", S)


if __name__ == "__main__":
    sys.exit(main())
