"""Generates tests/golden/*.npz with the CPU oracle.  The reference itself (Rust) cannot be run in this image and ships no
golden trajectories, so these fixtures pin the ORACLE's output: test_golden.py re-checks the oracle against them on every
run (guards against accidental changes of the restatement or of the RNG specification) and the GPU tests compare the CUDA
path against the same files.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nuts_rs_b200 import _abi  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (kind, model kwargs, N, d, settings overrides, n_draws, x0)
    "c1_iso_mu3_maxdepth3": (_abi.NUTS_LOGP_GAUSS_ISO, dict(mu=3.0), 4, 10, dict(num_tune=1000, maxdepth=3), 40, "const3.5"),
    "iso_mu05_noadapt": (_abi.NUTS_LOGP_GAUSS_ISO, dict(mu=0.5), 4, 10, dict(num_tune=0, maxdepth=6), 60, "normal"),
    "diag_d100_noadapt": (_abi.NUTS_LOGP_GAUSS_DIAG, dict(mu=0.5, sigma="logspace"), 3, 100, dict(num_tune=0, maxdepth=6), 30, "normal"),
    "rank1_d20_noadapt": (_abi.NUTS_LOGP_GAUSS_RANK1, dict(mu=0.0, rank1_scale=0.5), 3, 20, dict(num_tune=0, maxdepth=6), 30, "normal"),
    "funnel_d10": (_abi.NUTS_LOGP_FUNNEL, dict(funnel_scale=3.0), 4, 10, dict(num_tune=20, maxdepth=8), 12, "normal"),
    # low-rank mass matrix: a fixed rank-2 transformation per chain installed after set_position (see lowrank_transform())
    "lowrank_rank1_d20_noadapt": (_abi.NUTS_LOGP_GAUSS_RANK1, dict(mu=0.0, rank1_scale=0.5), 3, 20, dict(num_tune=0, maxdepth=6), 30, "normal"),
}


def lowrank_transform(N, d, r=2):
    """the transformation of the low-rank golden case: (stds, mean, vals [N, r], vecs [N, r, d], mean_low_rank).  Built from uniform
    random numbers, +-1 patterns and one correctly rounded square root only, so that it is the same bits on every machine."""
    assert d % 4 == 0 and r == 2 and N <= 3
    rng = np.random.default_rng(321)
    stds, mean = 0.75 + 0.5 * rng.random((N, d)), 0.2 * (rng.random((N, d)) - 0.5)
    i = np.arange(d)
    patterns = [np.ones(d), np.where(i % 2 == 0, 1.0, -1.0), np.where(i % 4 < 2, 1.0, -1.0)]  # mutually orthogonal for d % 4 == 0
    pairs = [(0, 1), (0, 2), (1, 2)]
    vecs = np.stack([np.stack([patterns[a], patterns[b]]) / np.sqrt(float(d)) for a, b in pairs[:N]])
    return stds, mean, 0.5 + 2.0 * rng.random((N, r)), vecs, 0.1 * (rng.random((N, d)) - 0.5)


def build(name):
    kind, mk, N, d, so, n_draws, x0kind = CASES[name]
    mk = dict(mk)
    if mk.get("sigma") == "logspace":
        mk["sigma"] = np.exp(np.linspace(-1, 1, d))
    s = _abi.default_settings()
    for k, v in so.items():
        setattr(s, k, v)
    x0 = np.full((N, d), 3.5) if x0kind == "const3.5" else np.random.default_rng(123).normal(size=(N, d))
    samp = O.Sampler(O.Model(kind, d, **mk), s, seed=2024, nchains=N)
    status = samp.set_position(x0)
    if name.startswith("lowrank"):
        assert samp.set_lowrank_transform(*lowrank_transform(N, d)).all()
    st0 = samp.state()
    draws, stats = samp.draw(n_draws)
    stats.pop("_total_leapfrogs")
    return dict(x0=x0, status=status, step_size0=st0["step_size"], stds0=st0["stds"], mean0=st0["mean"], draws=draws,
                **{"stat_" + k: v for k, v in stats.items()})


# KineticEnergyKind::ExactNormal / Microcanonical (Tier 2): a short trajectory of Hamiltonian::leapfrog per kind, and whole NUTS draws
# of the oracle with trajectory_kind set (the device runs these kinds at Tier 1 / Tier 2 only)
KINETIC_D, KINETIC_N, KINETIC_STEPS = 37, 3, 6


def kinetic_inputs():
    rng = np.random.default_rng(77)
    d, N = KINETIC_D, KINETIC_N
    return dict(sigma=np.exp(np.linspace(-1, 1, d)), stds=0.75 + 0.5 * rng.random((N, d)), mean=0.2 * (rng.random((N, d)) - 0.5),
                x0=rng.normal(size=(N, d)), eps=0.05 + 0.02 * rng.random(N))


def build_kinetic():
    inp = kinetic_inputs()
    d, N = KINETIC_D, KINETIC_N
    om = O.Model(_abi.NUTS_LOGP_GAUSS_DIAG, d, mu=0.5, sigma=inp["sigma"])
    out = {}
    for kind, tag in ((_abi.NUTS_KINETIC_EXACT_NORMAL, "exact"), (_abi.NUTS_KINETIC_MICROCANONICAL, "micro")):
        z, v, sc = np.zeros((KINETIC_STEPS, N, d)), np.zeros((KINETIC_STEPS, N, d)), np.zeros((KINETIC_STEPS, N, 3))
        for c in range(N):
            h = O.Hamiltonian(om)
            h.set_kinetic_energy_kind(kind)
            h.set_transform(inp["stds"][c], inp["mean"][c])
            p, st = h.init_state(inp["x0"][c])
            h.initialize_trajectory(p, True, 9, c + 1, 3)
            eps = inp["eps"][c] / (np.sqrt(float(d)) if kind == _abi.NUTS_KINETIC_MICROCANONICAL else 1.0)
            for k in range(KINETIC_STEPS):
                p, st, ee = h.leapfrog(p, eps, 1 if c != 1 else -1)
                z[k, c], v[k, c] = p.vec(p.Z), p.vec(p.V)
                s_ = p.scalars()
                sc[k, c] = s_["logp"], s_["kinetic_energy"], ee
        out.update({tag + "_z": z, tag + "_v": v, tag + "_scalars": sc})
    s = _abi.default_settings()
    s.num_tune, s.maxdepth, s.trajectory_kind = 30, 6, _abi.NUTS_KINETIC_EXACT_NORMAL
    samp = O.Sampler(O.Model(_abi.NUTS_LOGP_GAUSS_ISO, 10, mu=3.0), s, seed=2024, nchains=4)
    samp.set_position(np.full((4, 10), 3.5))
    draws, stats = samp.draw(40)
    out.update(nuts_exact_draws=draws, nuts_exact_depth=stats["depth"], nuts_exact_step_size=stats["step_size"],
               nuts_exact_energy_error=stats["energy_error"])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "kinetic_kinds.npz"), **build_kinetic())
    print("wrote kinetic_kinds")
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
    # RNG specification vectors
    normals, ctr = O.fill_normal(42, 1, 0, 64)
    unif = np.array([O.lib().orc_next_f64(42, 3, c) for c in range(16)])
    bools = np.array([O.lib().orc_next_bool(42, 3, c) for c in range(64)], dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "rng_spec.npz"), normals=normals, uniforms=unif, bools=bools)
    print("wrote rng_spec")
