"""Sharding of the control solve over the GPUs of one node (one process per GPU, torch.distributed).

The path partitions into independent units at three granularities:
  * flat candidates and whole episodes: contiguous block split, no data-path collective (`shard_range`);
  * the laps of ONE controller step: each rank solves a subset of the active laps' lambda chains, then the
    per-lap cost vectors (K doubles + K indices per lap, <= 4 KB at L=8, K=64) are all-gathered and every rank
    applies the reference's lexicographic winner rule locally (`gather_lap_costs`, `lexicographic_winner`).
    NCCL over NVLink when the tensors are CUDA tensors; gloo works for the CPU tests.
"""
import numpy as np


def shard_range(n_items, rank, world_size):
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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


class SplitStep:
    """One controller step whose laps are solved by `world_size` handles, one per rank / GPU (include/i2lqr.h,
    i2lqr_split_*).  Every rank holds the same safe set and calls `solve_step` with the same state; rank r runs
    the lambda chains of the active laps with position % world_size == r, the packed lap rows (costs, lambdas,
    iteration counts and the lap's best trajectory: 1 + 3K + 6N + 4 doubles per lap) are all-gathered once per
    outer iteration -- `ncclAllGather` over NVLink when the group's backend is NCCL -- and every rank applies
    the reference's winner rule to the same table, so all ranks return bit-identical results.

    torch supplies the device buffers, the stream and the collective; the kernels are the library's own.
    """

    def __init__(self, engine, rank, world_size, n_episodes=1, group=None, device=None):
        import torch

        self.torch = torch
        self.eng, self.rank, self.world, self.group = engine, int(rank), int(world_size), group
        self.device = torch.device("cuda", engine.device) if device is None else device
        L = engine.params.num_ss_iter
        self.row = engine.split_row_doubles()
        n = n_episodes * L * self.row
        self.n_episodes = n_episodes
        self.rows = torch.zeros(n, dtype=torch.float64, device=self.device)
        self.all_rows = torch.zeros(self.world * n, dtype=torch.float64, device=self.device)
        # a stream of our own: the library takes NULL as "the handle's stream", which torch's default stream is not
        self.stream = torch.cuda.Stream(self.device)
        torch.cuda.synchronize(self.device)

    def exchange(self):
        """rows of every rank -> all_rows, laid out (world, E, L, row)."""
        import torch.distributed as dist

        dist.all_gather_into_tensor(self.all_rows, self.rows, group=self.group)

    def solve_step(self, x):
        torch = self.torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            st = self.stream.cuda_stream
            E = self.eng.split_begin(x, st)
            assert E == self.n_episodes
            for it in range(3):  # calc_input's outer iterations, utils/base.py:384
                self.eng.split_solve(E, it, self.rank, self.world, self.rows.data_ptr(), st)
                self.exchange()
                self.eng.split_merge(E, it, self.rank, self.world, self.all_rows.data_ptr(), st)
            return self.eng.split_end(E, st)


def split_step_local(engines, x):
    """The split step with all parts on ONE device and the all-gather replaced by a concatenation: what the
    single-GPU parity test runs (same kernels, same row layout, no NCCL).  Returns every part's outputs."""
    import torch

    n_parts = len(engines)
    dev = torch.device("cuda", engines[0].device)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 4)
    E, L = x.shape[0], engines[0].params.num_ss_iter
    row = engines[0].split_row_doubles()
    rows = [torch.zeros(E * L * row, dtype=torch.float64, device=dev) for _ in engines]
    stream = torch.cuda.Stream(dev)  # not the default stream: the library reads NULL as "the handle's own stream"
    torch.cuda.synchronize(dev)
    with torch.cuda.device(dev), torch.cuda.stream(stream):
        st = stream.cuda_stream
        for eng in engines:
            eng.split_begin(x, st)
        for it in range(3):
            for part, eng in enumerate(engines):
                eng.split_solve(E, it, part, n_parts, rows[part].data_ptr(), st)
            all_rows = torch.cat(rows)
            for part, eng in enumerate(engines):
                eng.split_merge(E, it, part, n_parts, all_rows.data_ptr(), st)
            stream.synchronize()  # all_rows must outlive the merges that read it
        return [eng.split_end(E, st) for eng in engines]
