"""Deterministic integer 'policy' shared by the caller-parity fixtures and tests.

The action is an exact function of (observation, metadata, mask flag) in int64 / float64 arithmetic on small
integers — identical on CPU and GPU and for a batch or a single sample — so the reference's per-agent caller loops
(ppo.py:31-131, utils.py:500-573, :728-814) and the batched GPU adapters can be compared bit for bit.  With
``targets`` it walks at the opponent flag (home when carrying, read from metadata slot 7) three steps out of four,
so captures, tags, adjusted and terminal rewards occur in the fixtures; the rest is a hash of everything it sees.

It has the reference ``Agent`` surface (agent_network.py:42-81): ``get_action`` returns a Python int for a single
sample like ``Agent.get_action`` (``action.item()``, :58); ``get_action_and_value`` is batch-safe.
"""
import torch


class HashPolicy(torch.nn.Module):
    def __init__(self, n_obs, n_meta, salt, targets=None):
        """targets: ((opp_flag_row, col), (own_flag_row, col)) in the policy's frame, or None for the pure hash."""
        super().__init__()
        g = torch.Generator().manual_seed(salt)
        self.register_buffer("w1", torch.randint(1, 97, (n_obs,), generator=g).double())
        self.register_buffer("w2", torch.randint(1, 97, (n_meta,), generator=g).double())
        self.targets = targets

    def _act(self, grid, meta, use_action_mask):
        h = (grid.double().flatten(1) @ self.w1 + (meta.double() * 4096).round() @ self.w2).long()
        n = torch.where(use_action_mask.reshape(-1) == 1, 5, 9)
        rnd = h % n
        if self.targets is None:
            return rnd
        G = grid.shape[-1]
        idx = grid[:, 0].flatten(1).argmax(1)                 # own position plane
        r, c = idx // G, idx % G
        carrying = meta[:, 7] > 0.5
        (orow, ocol), (hrow, hcol) = self.targets
        dr = torch.where(carrying, hrow, orow) - r
        dc = torch.where(carrying, hcol, ocol) - c
        vert = torch.where(dr < 0, 0, 1)
        horiz = torch.where(dc > 0, 2, 3)
        use_vert = (dr != 0) & ((dc == 0) | ((h // 16) % 2 == 0))
        pref = torch.where(use_vert, vert, horiz)
        pref = torch.where((dr == 0) & (dc == 0), 4, pref)
        second = (n == 9) & (pref < 4) & ((h // 128) % 3 == 0)   # vault / place-block action set of types 2 and 3
        pref = torch.where(second, pref + 5, pref)
        explore = (h // 32) % 4 == 0
        return torch.where(explore, rnd, pref)

    def get_action(self, grid, meta, use_action_mask):
        a = self._act(grid, meta, use_action_mask)
        return a.item() if a.numel() == 1 else a

    def get_action_and_value(self, grid, meta, use_action_mask, action=None):
        a = self._act(grid, meta, use_action_mask)
        return a, -a.float() / 8, torch.zeros_like(a, dtype=torch.float32), (a.float() * 0.5).unsqueeze(1)


def flag_targets(env_or_ce):
    """((opponent flag), (own flag)) in the frame every policy sees (team 1 looks at the flipped grid)."""
    f = env_or_ce.FLAG_POSITIONS
    return (tuple(int(x) for x in f[1]), tuple(int(x) for x in f[0]))
