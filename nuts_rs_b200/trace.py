"""Draws + sampler statistics in the reference's trace schema.

The reference hands every draw to a storage backend that lays the trace out as groups `posterior` / `sample_stats` of arrays
with dimensions `[chain, draw, ...]` (src/storage/core.rs:12-77; tests/sample_normal.rs:262-267 reads
`/sample_stats/diverging` with dimension names ["chain", "draw"]).  The per-draw statistics and their names come from
`NutsStats` (src/chain.rs:215-231: depth, maxdepth_reached, chain, draw), `HamiltonianStats` (src/dynamics/
transformed_hamiltonian.rs:499-505: step_size), the step-size `Stats` (src/stepsize/adapt.rs:274-281: step_size_bar,
mean_tree_accept, mean_tree_accept_sym, n_steps, max_energy_error), `PointStats` (transformed_hamiltonian.rs:96-112:
index_in_trajectory, logp, energy, energy_error, fisher_distance), `DivergenceStats` (src/dynamics/hamiltonian.rs:39-56:
diverging) and the adaptation strategy's `tuning` flag (src/adapt_strategy.rs:246,280).

`nuts_draw` returns the same quantities draw-major (`[draw, chain]`, one kernel launch for all chains); this module only
re-labels and transposes them - it is host-side glue, not part of the device path."""
import numpy as np

#: statistic name in the reference's trace  ->  key of the nuts_draw statistics (include/nuts_b200.h nuts_stats_t)
SAMPLE_STATS = {
    "depth": "depth",
    "maxdepth_reached": "maxdepth_reached",
    "index_in_trajectory": "index_in_trajectory",
    "logp": "logp",
    "energy": "energy",
    "energy_error": "energy_error",
    "fisher_distance": "fisher_distance",
    "diverging": "diverging",
    "step_size": "step_size",
    "step_size_bar": "step_size_bar",
    "mean_tree_accept": "mean_tree_accept",
    "mean_tree_accept_sym": "mean_tree_accept_sym",
    "n_steps": "n_steps",
    "max_energy_error": "max_energy_error",
    "tuning": "tuning",
}
BOOL_STATS = ("maxdepth_reached", "diverging", "tuning")


def to_trace(draws, stats, chain_offset=0, draw_offset=0, parameter_name="unconstrained_draw"):
    """`draws` [n_draws, nchains, dim] and `stats` {name: [n_draws, nchains]} (what Sampler.draw returns) ->
    {"posterior": {...}, "sample_stats": {...}, "dims": {...}} with every array laid out `[chain, draw, ...]`.

    `chain_offset` is the global id of the first chain (multi-GPU shards), `draw_offset` the number of draws the sampler had
    produced before this call: together they give the reference's `chain` and `draw` statistics (src/sampler.rs:165-174)."""
    out = {"posterior": {}, "sample_stats": {}, "dims": {}}
    n_draws, nchains = next(iter(stats.values())).shape if stats else draws.shape[:2]
    if draws is not None:
        out["posterior"][parameter_name] = np.ascontiguousarray(np.transpose(draws, (1, 0, 2)))
        out["dims"]["posterior/" + parameter_name] = ["chain", "draw", "unconstrained_parameter"]
    for name, key in SAMPLE_STATS.items():
        if key not in stats:
            continue
        a = np.ascontiguousarray(stats[key].T)
        if name in BOOL_STATS:
            a = a.astype(bool)
        out["sample_stats"][name] = a
        out["dims"]["sample_stats/" + name] = ["chain", "draw"]
    chain = np.arange(chain_offset, chain_offset + nchains, dtype=np.uint64)
    draw = np.arange(draw_offset, draw_offset + n_draws, dtype=np.uint64)
    out["sample_stats"]["chain"] = np.ascontiguousarray(np.broadcast_to(chain[:, None], (nchains, n_draws)))
    out["sample_stats"]["draw"] = np.ascontiguousarray(np.broadcast_to(draw[None, :], (nchains, n_draws)))
    out["dims"]["sample_stats/chain"] = ["chain", "draw"]
    out["dims"]["sample_stats/draw"] = ["chain", "draw"]
    return out


def concat(traces, axis):
    """Join traces along "chain" (shards of a multi-GPU run, in rank order) or "draw" (consecutive nuts_draw calls)."""
    ax = {"chain": 0, "draw": 1}[axis]
    out = {"posterior": {}, "sample_stats": {}, "dims": dict(traces[0]["dims"])}
    for group in ("posterior", "sample_stats"):
        for k in traces[0][group]:
            out[group][k] = np.concatenate([t[group][k] for t in traces], axis=ax)
    return out


def save_npz(path, trace):
    """One .npz with keys `posterior/<name>` and `sample_stats/<name>` (+ `dims/<group>/<name>`: the dimension names)."""
    flat = {}
    for group in ("posterior", "sample_stats"):
        for k, v in trace[group].items():
            flat[f"{group}/{k}"] = v
    for k, v in trace["dims"].items():
        flat["dims/" + k] = np.array(v)
    np.savez(path, **flat)


def load_npz(path):
    z = np.load(path)
    out = {"posterior": {}, "sample_stats": {}, "dims": {}}
    for k in z.files:
        group, name = k.split("/", 1)
        if group == "dims":
            out["dims"][name] = [str(x) for x in z[k]]
        else:
            out[group][name] = z[k]
    return out
