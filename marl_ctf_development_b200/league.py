"""Duel batches for league evaluation on GPU envs.

The reference evaluates a league by submitting ``number_of_duels`` Ray tasks per (agent, opponent) pair, each a
full ``utils.duel`` on a pickled env copy (league_training.py:368-459, 573-648).  Here every duel is one env of a
batch: pair p owns the envs [p*D, (p+1)*D) and all pairs advance with one step-kernel launch per step; only the
policy forwards are per pair.  Results come back in the reference's dict formats.
"""
from __future__ import annotations

from collections import defaultdict

import torch

from .env import GridworldCtfGPU, metrics_dict


@torch.no_grad()
def duel_pairs(env_config: dict, pairs, duels_per_pair: int, max_steps: int = 256, device=None, seed: int = 0,
               env_id_base: int = 0, collect_metrics: bool = False):
    """Plays ``duels_per_pair`` duels (utils.py:500-573) for every (agent, opponent) in ``pairs`` at once.

    Returns (results [P, D] int64 of +1/0/-1 from team 0's view, metrics) where metrics is a list of P
    ``env.metrics``-style dicts summed over the pair's duels (None unless collect_metrics).
    """
    P, D = len(pairs), int(duels_per_pair)
    env = GridworldCtfGPU(**env_config, num_envs=P * D, device=device, seed=seed, env_id_base=env_id_base,
                          reverse_team1_actions=True, stats="counters" if collect_metrics else "none")
    B, N, dev = env.num_envs, env.N_AGENTS, env.device
    C, G, M = env.n_channels, env.GRID_SIZE, env.meta_size
    teams = [torch.tensor([i for i in range(N) if env.AGENT_TEAMS[i] == t], device=dev) for t in (0, 1)]
    obs, meta, _ = env.reset()
    actions = torch.empty((B, N), dtype=torch.uint8, device=dev)
    step_count = 0
    while True:
        step_count += 1
        for p, pair in enumerate(pairs):
            sl = slice(p * D, (p + 1) * D)
            for idx, pol in zip(teams, pair):
                k = idx.numel()
                if k == 0:
                    continue
                a = pol.get_action(
                    obs[sl][:, idx].reshape(D * k, C, G, G).float(), meta[sl][:, idx].reshape(D * k, M),
                    env.use_action_mask[idx].unsqueeze(0).expand(D, k).reshape(D * k),
                )
                actions[sl, idx] = a.reshape(D, k).to(torch.uint8)
        obs, meta, _, _, _ = env.step(actions)
        if step_count > max_steps or step_count >= env.GAME_STEPS:
            break
    caps = env.flag_captures()
    results = torch.sign(caps[:, 0] - caps[:, 1]).reshape(P, D)
    metrics = None
    if collect_metrics:
        c = env.counters()
        per_pair = c.reshape(P, D, *c.shape[1:]).sum(1).cpu().numpy()
        metrics = [metrics_dict(env.ce, per_pair[p]) for p in range(P)]
    env.close()
    return results, metrics


def _matrices(results, labels, number_of_duels):
    winrate, draw = defaultdict(int), defaultdict(int)
    res = results.cpu()
    for p, (a_label, o_label) in enumerate(labels):
        for d in range(number_of_duels):
            r = int(res[p, d])
            winrate[(a_label, o_label)] += (r == 1) / number_of_duels   # league_training.py:391-392 / :446-447
            draw[(a_label, o_label)] += (r == 0) / number_of_duels
    for a_label, o_label in list(winrate.keys()):                        # :395-397 / :450-452
        winrate[(o_label, a_label)] = 1 - winrate[(a_label, o_label)] - draw[(a_label, o_label)]
    return winrate, draw


def winrate_matrix_non_symmetric(env_config, agents_t1, agents_t2, number_of_duels, **kw):
    """league_training.py:427-455: every team-1 agent against every team-2 agent; keys ('0_i', '1_j') and inverses."""
    pairs, labels = [], []
    for i, agent in enumerate(agents_t1):
        for j, opponent in enumerate(agents_t2):
            pairs.append((agent, opponent))
            labels.append((f"0_{i}", f"1_{j}"))
    results, _ = duel_pairs(env_config, pairs, number_of_duels, **kw)
    return _matrices(results, labels, number_of_duels)[0]


def winrate_matrix_symmetric(env_config, agents, number_of_duels, **kw):
    """league_training.py:368-400: pairs with agent_idx > opponent_idx; keys ('0_i', '0_j') and inverses."""
    pairs, labels = [], []
    for i, agent in enumerate(agents):
        for j, opponent in enumerate(agents):
            if i > j:
                pairs.append((agent, opponent))
                labels.append((f"0_{i}", f"0_{j}"))
    if not pairs:
        return defaultdict(int)
    results, _ = duel_pairs(env_config, pairs, number_of_duels, **kw)
    return _matrices(results, labels, number_of_duels)[0]
