import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
np.set_printoptions(linewidth=250, precision=4)
from nuts_rs_b200 import _abi, lib
d = int(sys.argv[1]); N = int(sys.argv[2]); tune = int(sys.argv[3]); wide = int(sys.argv[4])
sig = 10 ** np.linspace(-3, 3, d) if wide else np.exp(np.linspace(-1, 1, d))
out = {}
for eng in sys.argv[5:]:
    os.environ["NUTS_B200_ENGINE"] = eng
    m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.0, sigma=sig)
    s = lib.Sampler(m, lib.DiagNutsSettings(num_tune=tune, maxdepth=int(os.environ.get("MAXDEPTH", 6))), seed=42)
    st = s.set_position(np.random.default_rng(42).normal(size=(N, d)))
    state = s.state()
    draws, stats = s.draw(12)
    state = s.state()
    out[eng] = (state, draws, stats)
    print(eng, "status", st[:6], "eps", state["step_size"][:6], "\n  depth c0", stats["depth"][:, 0], "div c0", stats["diverging"][:, 0], "\n  eps c0", stats["step_size"][:, 0], "\n  mean depth per draw", stats["depth"].mean(axis=1), "div frac", stats["diverging"].mean(axis=1), "\n  chains with any div after draw 5:", np.nonzero(stats["diverging"][5:].any(axis=0))[0][:20])
    s.close(); m.close()
a, b = [out[e] for e in sys.argv[5:7]]
print("stds rel diff per chain", np.abs(a[0]["stds"] / b[0]["stds"] - 1).max(axis=1)[:6], "mean diff", np.abs(a[0]["mean"] - b[0]["mean"]).max(axis=1)[:6])
bad = np.argwhere(np.abs(a[0]["stds"] / b[0]["stds"] - 1) > 1e-9)
print("bad stds entries", bad[:8].tolist(), len(bad))
