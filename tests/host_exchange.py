"""Host-side model of the one exchange on the path: when the laps of ONE controller step are
solved by several ranks, the per-lap cost rows are all-gathered and every rank applies the reference's winner
rule.  The product does this on the device (`i2lqr_split_*`, `winner_kernel`, NCCL through `dist.SplitStep`);
these NumPy / torch.distributed stand-ins exist so that the rule and the exchange pattern can be tested on CPU
with the gloo backend (tests/test_host.py).  Test scaffolding, not product code."""
import numpy as np


def lap_owner(n_active_laps, world_size):
    """Round-robin owner rank of each active lap of one controller step."""
    return [l % world_size for l in range(n_active_laps)]


def lexicographic_winner(costs, counts):
    """The reference's winner rule (iterative_ilqr/utils/base.py:462-465): `min(cost_list)` is Python's
    lexicographic list comparison over per-lap cost vectors, then the first arg-min inside that lap.

    costs (L,K) float array, counts (L,) valid entries per lap.  Returns (lap, rank).
    """
    best = 0
    for l in range(1, len(counts)):
        a, b = costs[l, : counts[l]], costs[best, : counts[best]]
        n = min(len(a), len(b))
        less = len(a) < len(b)
        for i in range(n):
            if a[i] != b[i]:
                less = a[i] < b[i]
                break
        if less:
            best = l
    row = costs[best, : counts[best]]
    rank = 0
    for r in range(1, len(row)):
        if row[r] < row[rank]:
            rank = r
    return best, rank


def gather_lap_costs(local_costs, local_counts, owners, rank, group=None, device=None):
    """All-gather of the per-lap (cost, count) rows when one step's laps span ranks.

    local_costs (L,K): rows of laps this rank owns are filled, others ignored.  Every rank returns the full
    (L,K) cost matrix and (L,) counts.  One all_gather of L*(K+1) doubles per outer iteration.
    """
    import torch
    import torch.distributed as dist

    L, K = local_costs.shape
    payload = np.zeros((L, K + 1))
    payload[:, :K] = local_costs
    payload[:, K] = local_counts
    t = torch.from_numpy(payload)
    if device is not None:
        t = t.to(device)
    world = dist.get_world_size(group)
    bucket = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bucket, t, group=group)
    full = np.zeros((L, K + 1))
    for l in range(L):
        full[l] = bucket[owners[l]][l].cpu().numpy()
    return full[:, :K], full[:, K].astype(np.int64)
