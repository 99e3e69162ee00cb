"""One eager training step on the 10k-bus feeder (BASELINE config 5, large-graph path) for `ncu --profile-from-start off` launch lists.
usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file X python tools/feeder_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch
from dss2 import synth
from dss2.trainer import GraphedTrainer, default_spec

REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
grid = synth.replicate_feeder(synth.load_grid("ober_sub"), 143)
store = synth.synthetic_store(grid, 64, seed=7, device="cuda")
tr = GraphedTrainer(store, 32, spec=default_spec(), reg_coefs=REG, seed=0, use_cuda_graph=False)
ids = torch.arange(32, device="cuda")
for _ in range(2):
    tr.step(ids)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(ids)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(tr.loss))
