"""Tier-2 fused leapfrog (plane kernel k_leapfrog) at an HBM-resident size, for an ncu timing of the kernel itself:
16384 chains x 1000 dims = 131 MB per plane, 10 planes touched per step (read z, v, grad_z, sigma, mu; write z', v', x', grad_x', grad_z')."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nuts_rs_b200 import _abi, lib
N, d = 16384, 1000
m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
m.set_transform(np.exp(np.linspace(-0.5, 0.5, d)), 0.1)
p, st = m.init_state(np.random.default_rng(0).normal(size=(N, d)))
m.initialize_trajectory(p, True, 42, 0, 0)
q = lib.Point(m)
for _ in range(6):
    q, status, ee = m.leapfrog(p, 0.1, out=q)
    p, q = q, p
print("ok", int((status != 0).sum()), float(np.abs(ee).max()), "algorithmic bytes per launch", 80 * d * N)
