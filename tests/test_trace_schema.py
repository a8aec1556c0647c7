"""The trace schema glue (nuts_rs_b200/trace.py) against the reference's names and layout, with the CPU oracle as the draw
source: tests/sample_normal.rs:262-267 (dimension names ["chain", "draw"]), src/sampler.rs:1662-1692 (the `draw` index counts
up, `tuning` flips after num_tune), src/chain.rs:215-231 / src/stepsize/adapt.rs:274-281 (statistic names)."""
import numpy as np

from nuts_rs_b200 import _abi, trace
from oracle import oracle as O


def _oracle_run(n_tune, n_draws, nchains=3, d=5):
    from nuts_rs_b200 import lib  # only for the settings struct defaults (no device call)

    s = _abi.default_settings()
    s.num_tune = n_tune
    s.maxdepth = 4
    m = O.Model(_abi.NUTS_LOGP_GAUSS_ISO, d, mu=3.0)
    samp = O.Sampler(m, s, seed=5, nchains=nchains)
    samp.set_position(np.full((nchains, d), 3.5))
    return samp.draw(n_draws)


def test_names_dims_and_layout(tmp_path):
    draws, stats = _oracle_run(6, 10)
    t = trace.to_trace(draws, stats, chain_offset=8, draw_offset=0)
    want = {"depth", "maxdepth_reached", "chain", "draw", "step_size", "step_size_bar", "mean_tree_accept", "mean_tree_accept_sym",
            "n_steps", "max_energy_error", "index_in_trajectory", "logp", "energy", "energy_error", "fisher_distance", "diverging",
            "tuning"}
    assert want <= set(t["sample_stats"])
    for k in want:
        assert t["sample_stats"][k].shape == (3, 10)
        assert t["dims"]["sample_stats/" + k] == ["chain", "draw"]
    assert t["posterior"]["unconstrained_draw"].shape == (3, 10, 5)
    assert t["dims"]["posterior/unconstrained_draw"] == ["chain", "draw", "unconstrained_parameter"]
    # [chain, draw] is the transpose of nuts_draw's [draw, chain]
    assert np.array_equal(t["posterior"]["unconstrained_draw"][1, 7], draws[7, 1])
    assert np.array_equal(t["sample_stats"]["logp"][2], stats["logp"][:, 2])
    # Progress bookkeeping (src/sampler.rs:1662-1692): draw counts up from 0, chain is the global id, tuning flips after num_tune
    assert (t["sample_stats"]["draw"] == np.arange(10)[None, :]).all()
    assert (t["sample_stats"]["chain"] == np.array([8, 9, 10])[:, None]).all()
    assert t["sample_stats"]["tuning"].dtype == bool
    assert t["sample_stats"]["tuning"][:, :6].all() and not t["sample_stats"]["tuning"][:, 6:].any()
    assert t["sample_stats"]["diverging"].dtype == bool and not t["sample_stats"]["diverging"].any()
    # npz round trip
    p = tmp_path / "trace.npz"
    trace.save_npz(p, t)
    back = trace.load_npz(p)
    for g in ("posterior", "sample_stats"):
        for k, v in t[g].items():
            assert np.array_equal(back[g][k], v, equal_nan=True)
    assert back["dims"] == t["dims"]


def test_concat_chains_and_draws():
    d1, s1 = _oracle_run(0, 4)
    d2, s2 = _oracle_run(0, 4)
    a = trace.to_trace(d1, s1, chain_offset=0)
    b = trace.to_trace(d2, s2, chain_offset=3)
    c = trace.concat([a, b], "chain")
    assert c["posterior"]["unconstrained_draw"].shape == (6, 4, 5)
    assert (c["sample_stats"]["chain"][:, 0] == np.arange(6)).all()
    e = trace.concat([a, trace.to_trace(d2, s2, draw_offset=4)], "draw")
    assert (e["sample_stats"]["draw"][0] == np.arange(8)).all()
