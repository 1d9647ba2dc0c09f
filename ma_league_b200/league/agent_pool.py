"""League parameter exchange on the device.

The reference moves agent parameters between league instances as pickled, cloned `state_dict`s through
multiprocessing Queues to a central `AgentPoolInstance` process and back (league/processes/agent_pool_instance.py:
13-128, league/utils/commands.py:22-70, league_experiment_process.py:63-83), once per matchup (every `play_time_mins`).
With one league instance per GPU of an NVSwitch box the pool becomes one `[world, P]` device tensor per rank:

    pool.sync()                 every rank contributes the flat fp32 buffer of its home agent; one all-gather over
                                NCCL / NVLink (8 x 119 KB at 5v5) replaces the Queue round trips and the Barrier
    pool.state_dict(idx)        the pooled agent `idx` as an OrderedDict with the reference's keys (fc1.weight, ...,
                                fc2.bias): what `AgentParamsGetCommand` returned; values are views of the pool tensor
    pool.load_into(mac, idx)    `mac.load_state_dict(agent=pool.state_dict(idx))` as one device-to-device copy of the
                                flat buffer (sp_ma_experiment.py:27-29 loads the opponent this way)

Matchmaking (who plays whom) stays where it is in the reference; only the parameter transport changes.  `state_dict`
keys, shapes and order are those of `mac.agent.state_dict()`, so checkpoints and the reference's commands interoperate.
"""
from collections import OrderedDict

import torch as th
import torch.distributed as dist

from ..flat import ensure_flat, flat_views


class DeviceAgentPool:
    def __init__(self, mac, group=None):
        self.mac = mac
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        flat = ensure_flat(mac.agent)
        self.n_params = flat.numel()
        self.pool = th.zeros(self.world, self.n_params, dtype=th.float32, device=flat.device)
        self.n_syncs = 0

    def sync(self):
        """All ranks call this together (the reference's `_sync_barrier` point): pool[r] <- rank r's home agent."""
        flat = ensure_flat(self.mac.agent).detach()
        if self.world == 1:
            self.pool[0].copy_(flat)
        else:
            dist.all_gather(list(self.pool.unbind(0)), flat.contiguous(), group=self.group)
        self.n_syncs += 1
        return self.pool

    def state_dict(self, idx: int) -> OrderedDict:
        names = [k for k, _ in self.mac.agent.named_parameters()]
        views = flat_views(self.pool[idx], list(self.mac.agent.parameters()))
        return OrderedDict(zip(names, views))

    def load_into(self, mac, idx: int):
        """Make `mac` (e.g. the away controller) play with pooled agent `idx`."""
        dst = ensure_flat(mac.agent)
        if dst.numel() != self.n_params:
            raise ValueError("agent architectures differ: %d vs %d parameters" % (dst.numel(), self.n_params))
        with th.no_grad():
            dst.copy_(self.pool[idx].to(dst.device, non_blocking=True))
        return mac
