"""N > 1 host logic on CPU: world_size-2 gloo runs of the chain sharding + draw gather.  The per-rank "sampler"
is the CPU oracle with chain_id_offset = shard offset, which also proves the property the multi-GPU run relies on: a sharded
run reproduces the unsharded run chain by chain because RNG streams are keyed by the global chain id."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nuts_rs_b200 import _abi, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    for n in (1, 6, 7, 1024, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            r = sharding.all_shard_ranges(n, w)
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            for (o1, c1), (o2, _) in zip(r, r[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in r) - min(c for _, c in r) <= 1
    assert sharding.all_shard_ranges(65536, 8) == [(i * 8192, 8192) for i in range(8)]  # BASELINE config 5
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, num_chains, dim, n_draws, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    offset, count = sharding.shard_range(num_chains, world, rank)
    settings = _abi.default_settings()
    settings.num_tune = 10
    settings.maxdepth = 4
    x0 = np.random.default_rng(0).normal(size=(num_chains, dim))
    model = O.Model(_abi.NUTS_LOGP_GAUSS_ISO, dim, mu=0.5)
    samp = O.Sampler(model, settings, seed=7, nchains=count, chain_id_offset=offset)
    samp.set_position(x0[offset:offset + count])
    draws, stats = samp.draw(n_draws)
    total = stats.pop("_total_leapfrogs")
    full = sharding.gather_draws(torch.from_numpy(draws), num_chains)
    full_stats = sharding.gather_stats({"n_steps": stats["n_steps"], "diverging": stats["diverging"]}, num_chains)
    grand_total = sharding.total_leapfrogs(total)
    if rank == 0:
        np.save(os.path.join(out_dir, "draws.npy"), full.numpy())
        np.save(os.path.join(out_dir, "n_steps.npy"), full_stats["n_steps"])
        np.save(os.path.join(out_dir, "total.npy"), np.array([grand_total]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_chains", [6, 5])
def test_two_rank_gloo_gather_matches_unsharded_run(tmp_path, orc, num_chains):
    dim, n_draws, world = 8, 12, 2
    port = 29500 + (os.getpid() % 2000) + num_chains
    mp.spawn(_worker, args=(world, port, num_chains, dim, n_draws, str(tmp_path)), nprocs=world, join=True)
    settings = _abi.default_settings()
    settings.num_tune = 10
    settings.maxdepth = 4
    x0 = np.random.default_rng(0).normal(size=(num_chains, dim))
    model = orc.Model(_abi.NUTS_LOGP_GAUSS_ISO, dim, mu=0.5)
    samp = orc.Sampler(model, settings, seed=7, nchains=num_chains)
    samp.set_position(x0)
    draws, stats = samp.draw(n_draws)
    np.testing.assert_array_equal(np.load(tmp_path / "draws.npy"), draws)
    np.testing.assert_array_equal(np.load(tmp_path / "n_steps.npy"), stats["n_steps"])
    assert int(np.load(tmp_path / "total.npy")[0]) == int(stats["n_steps"].sum())
