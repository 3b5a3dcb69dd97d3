"""Sharding of the control solve over the GPUs of one node (one process per GPU, torch.distributed).

The path partitions into independent units at three granularities:
  * flat candidates and whole episodes: contiguous block split, no data-path collective (`shard_range`);
  * the laps of ONE controller step: each rank solves a subset of the active laps' lambda chains, then one packed
    row per lap (costs, lambdas, iteration counts, the lap's best trajectory) is all-gathered over NCCL and every
    rank runs the same device-side winner kernel on the same table (`SplitStep`).  (A host-side model of that
    exchange, used by the world-size-2 gloo test, lives in tests/host_exchange.py.)
"""
import numpy as np


def shard_range(n_items, rank, world_size):
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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
