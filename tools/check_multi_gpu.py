"""N-GPU check of the env-sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py

Every rank holds an env shard of a golden buffer and runs ``AgentPPO.update_net`` with data parallelism enabled
(GAE on the shard, all-reduced advantage sums, per-minibatch NCCL all-reduce of the flat gradient, replicated
clip + Adam).  Rank 0 also runs a single-GPU agent on the WHOLE buffer with the union of the ranks' minibatch
indices; the two must agree (rtol 1e-4) and all ranks must end with identical parameters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
import torch.distributed as dist

from tests import golden_utils as gu
from tests import gpu_utils as G


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    th.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=th.device(dev))
    case = os.environ.get("CHECK_CASE", "synth_s8_a2_64x64")   # 64 x 64 nets: the tcgen05 update kernel + peer exchange apply
    g = gu.load(case)
    h, n = g["buf.states"].shape[:2]
    if n % world:   # envs are independent columns: drop the remainder so that every rank owns the same number of envs
        n -= n % world
        assert n > 0, f"{case} has fewer envs than ranks"
        for k in list(g.keys()):
            if k.startswith("buf.") and g[k].ndim >= 2 and g[k].shape[0] == h:
                g[k] = np.ascontiguousarray(g[k][:, :n])
        g["buf.last_state"] = np.ascontiguousarray(g["buf.last_state"][:n])
    shard = n // world
    lo = rank * shard
    batch, updates = 16 * world, 3
    rng = np.random.default_rng(1234)
    local_ids = rng.integers(0, h * shard, (world, updates, batch // world))         # same on every rank
    env_l, t_l = local_ids // h, local_ids % h
    global_ids = np.concatenate([(env_l[r] + r * shard) * h + t_l[r] for r in range(world)], axis=1)  # [updates, batch]

    keys = ("states", "actions", "logprobs", "rewards", "undones", "unmasks")
    modes = ("peer", "peer_allreduce", "gather", "allreduce") if tuple(int(x) for x in g["dims"][4:]) == (64, 64) else ("gather", "allreduce")
    for mode in modes:
        run_mode(mode, g, rank, world, local, dev, h, n, shard, lo, batch, updates, local_ids, global_ids, keys)
    dist.barrier()
    dist.destroy_process_group()


def run_mode(mode, g, rank, world, local, dev, h, n, shard, lo, batch, updates, local_ids, global_ids, keys):
    agent = G.agent_from_golden(g, gpu_id=local, batch_size=batch, repeat_times=updates * batch / h + 1e-9, num_envs=shard)
    agent.enable_data_parallel()
    agent.sharded_mode = mode
    buf = [G.cuda(np.ascontiguousarray(g[f"buf.{k}"][:, lo:lo + shard]), dev) for k in keys]
    agent.last_state = G.cuda(g["buf.last_state"][lo:lo + shard], dev)
    agent._inject_ids = G.cuda(local_ids[rank], dev)
    res = agent.update_net(buf)
    assert agent.last_update_info["update_times"] == updates
    flat = th.cat([p.detach().reshape(-1) for p in list(agent.act.parameters()) + list(agent.cri.parameters())])
    gathered = [th.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        for r in range(1, world):
            assert th.equal(gathered[0], gathered[r]), f"rank {r} parameters differ from rank 0"
        single = G.agent_from_golden(g, gpu_id=local, batch_size=batch, repeat_times=updates * batch / h + 1e-9, num_envs=n)
        sbuf = [G.cuda(g[f"buf.{k}"], dev) for k in keys]
        single.last_state = G.cuda(g["buf.last_state"], dev)
        single._inject_ids = G.cuda(global_ids, dev)
        sres = single.update_net(sbuf)
        np.testing.assert_allclose(res, sres, rtol=1e-4, atol=1e-6)
        sflat = th.cat([p.detach().reshape(-1) for p in list(single.act.parameters()) + list(single.cri.parameters())])
        np.testing.assert_allclose(flat.cpu().numpy(), sflat.cpu().numpy(), rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(agent.last_update_info["adv_stats"][:2].cpu().numpy(),
                                   single.last_update_info["adv_stats"][:2].cpu().numpy(), rtol=1e-5)
        print(f"| multi-GPU check ok on {world} ranks [{mode}]: sharded == single-GPU, result {tuple(round(x, 6) for x in res)}")


if __name__ == "__main__":
    main()
