"""World-size-2 `gloo` test (CPU) of the env-sharded update's host-side algebra:

* advantage statistics: each rank reduces its env shard with the lattice on the GLOBAL env index; one all-reduce
  of the sums reproduces the reference's mean / std[::4, ::4] of the whole buffer;
* gradients: each rank's gradient sums over its share of the minibatch, scaled by 1/global_batch, all-reduced
  (sum), then clip + Adam replicated -- identical parameters on both ranks, equal to the single-process update on
  the union minibatch (the golden minted from the reference).
The per-rank arithmetic is the numpy oracle; the collectives are real torch.distributed gloo collectives across
two processes.  The CUDA kernels that fill the same buffers are covered by the `-m gpu` tests."""
import os
import tempfile

import numpy as np
import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_oracle as po
from tests import golden_utils as gu

CASES = ["synth_s8_a2_128x64", "discrete_s8_a4_128x64"]  # Gaussian policy; categorical policy (ragged shards: 9 envs)


def _worker(rank, world, init_file, out_dir, case):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    g = gu.load(case)
    discrete = case.startswith("discrete")
    hp = dict(gu.hyper_of(g), discrete=discrete)
    actor_of = gu.discrete_net_of if discrete else gu.net_of
    h, n = g["buf.states"].shape[:2]
    lo = rank * (n // world)
    shard = n - lo if rank == world - 1 else n // world

    # ---- (1) advantage statistics from shard sums
    sums = th.from_numpy(po.lattice_stat_sums(g["gae.advantages"][:, lo:lo + shard], lo))
    dist.all_reduce(sums)
    mean, std = po.stats_from_sums(sums.numpy(), h * n, ((h + 3) // 4) * ((n + 3) // 4))
    np.testing.assert_allclose([mean, std], [g["gae.adv_mean"], g["gae.adv_std"]], rtol=1e-5, atol=1e-7)

    # ---- (2) one minibatch: every rank takes the samples of the golden minibatch that live in its shard
    actor, critic = actor_of(g, "actor"), gu.net_of(g, "critic")
    opt_a, opt_c = po.new_adam_state(actor, not discrete), po.new_adam_state(critic, False)
    buffer = dict(states=g["buf.states"], actions=g["buf.actions"], unmasks=g["buf.unmasks"], logprobs=g["buf.logprobs"],
                  advantages=g["gae.adv_norm"], reward_sums=g["gae.reward_sums"])
    ids = g["update.ids"][0]
    _, env_of = po.split_ids(ids, h)
    mine = ids[(env_of >= lo) & (env_of < lo + shard)]
    global_batch = len(ids)
    scalars, ga, gc = po.ppo_minibatch_grads(actor, critic, po.gather_minibatch(buffer, mine), hp, denominator=global_batch)
    flat = th.from_numpy(np.concatenate([x.ravel() for x in ga + gc] + [np.array(scalars, dtype=np.float32)]))
    dist.all_reduce(flat)                                   # the one collective per minibatch
    flat = flat.numpy()
    off = 0
    ga_r, gc_r = [], []
    for src, dst in ((ga, ga_r), (gc, gc_r)):
        for x in src:
            dst.append(flat[off:off + x.size].reshape(x.shape).copy())
            off += x.size
    np.testing.assert_allclose(flat[off:off + 3], g["update.scalars"][0], rtol=1e-4, atol=1e-6)
    po.ppo_apply_grads(actor, critic, opt_a, opt_c, ga_r, gc_r, hp)
    for which, net in (("actor", actor), ("critic", critic)):
        ref_net = (actor_of if which == "actor" else gu.net_of)(g, f"update.after1.{which}")
        for mine_p, ref in zip(gu.flat_params(net), gu.flat_params(ref_net)):
            np.testing.assert_allclose(mine_p, ref, rtol=1e-4, atol=1e-6)
    np.save(os.path.join(out_dir, f"w_{rank}.npy"), np.concatenate([p.ravel() for p in gu.flat_params(actor)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", CASES)
def test_env_sharded_update_matches_single_process(case):
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_worker, args=(2, init_file, d, case), nprocs=2, join=True)
        w0, w1 = np.load(os.path.join(d, "w_0.npy")), np.load(os.path.join(d, "w_1.npy"))
        assert np.array_equal(w0, w1), "ranks must end with bit-identical parameters"
