import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nuts_rs_b200 import _abi, lib
d = int(sys.argv[1]); N = 2
m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
s = lib.Sampler(m, lib.DiagNutsSettings(num_tune=0, maxdepth=3), seed=7)
st = s.set_position(np.random.default_rng(1).normal(size=(N, d)))
print("status", st, "eps", s.state()["step_size"])
draws, stats = s.draw(2)
print("depth", stats["depth"].ravel(), "n_steps", stats["n_steps"].ravel(), "div", stats["diverging"].ravel(), "energy_error", stats["energy_error"].ravel(), "logp", stats["logp"].ravel())
