"""One process per GPU: read sharding and the rank-level reductions bench.py needs.

The `dtw` path shards by reads with the reference replicated, so there is no
collective on the data path.  torch.distributed is used only for the barrier around the timed region
and for reducing per-rank timings (max) and work counts (sum): NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items: int, rank: int, world: int):
    """contiguous, order-preserving split of n_items over world ranks (sizes differ by at most one)"""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class Ranks:
    def __init__(self, backend: str | None = None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = env_rank()
        self.device = device if device is not None else torch.device("cpu")
        self.owned = False
        self._gloo = None
        if self.world > 1 and not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = self.device
            dist.init_process_group(backend or "gloo", **kw)
            self.owned = True

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        if self.device.type == "cuda":
            self.torch.cuda.synchronize()

    def host_barrier(self):
        """barrier that waits on the CPU (a gloo group).  Ranks idling in an NCCL barrier keep a kernel spinning on
        their GPU; a process that uses that GPU meanwhile (bench.py's from-files run of the command line on all
        GPUs) is then time-sliced against it and runs at half speed (measured)."""
        if self.world > 1:
            if self._gloo is None:
                self._gloo = self.dist.new_group(backend="gloo")
            self.dist.barrier(group=self._gloo)

    def _reduce(self, values, op):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    def max(self, values):
        return self._reduce(values, self.dist.ReduceOp.MAX)

    def sum(self, values):
        return self._reduce(values, self.dist.ReduceOp.SUM)

    def gather_ordered(self, rows):
        """per-rank result rows (any picklable list) -> on rank 0 the concatenation in rank order, which
        is input order for shard_range() shards; None elsewhere"""
        if self.world == 1:
            return list(rows)
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(list(rows), out, dst=0)
        if self.rank != 0:
            return None
        return [r for part in out for r in part]

    def close(self):
        if self.owned:
            self.dist.destroy_process_group()


def job_throughput(cells_per_rank_sum: float, ms_per_step_max: float) -> float:
    """whole-job GCUPS: the cells all ranks processed in a step / the slowest rank's step time"""
    return cells_per_rank_sum / (ms_per_step_max * 1e-3) / 1e9
