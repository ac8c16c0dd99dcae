"""Multi-GPU plumbing for the hot path: one process per GPU, stereo pairs sharded over ranks.

The ops are independent per batch item, so the data path needs NO collective: each
rank runs the kernels on its own slice of the batch.  torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests) is used only for the rendezvous, the barriers that bracket a timed region and the max-over-ranks
reduction of device times -- mirroring the reference's DDP layout (torch_implementation.py:629,741,773-775).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class World:
    rank: int
    local_rank: int
    world_size: int
    backend: str | None      # None when running single-process

    @property
    def distributed(self) -> bool:
        return self.backend is not None

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def init_world(backend: str | None = None, force_group: bool = False) -> World:
    """Join the job described by RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun env).
    Single process (no env) -> a 1-rank world without a process group, unless force_group: then a 1-rank process group
    is created so that a single GPU runs exactly the code path of N GPUs (DDP, SyncBatchNorm conversion, exchange kernels)
    -- the N = 1 point of a scaling curve must be the same configuration as the N > 1 points."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if ws <= 1 and not force_group:
        return World(0, 0, 1, None)
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=ws, **kwargs)
    return World(rank, local, ws, backend)


def shard_range(total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of `total` stereo pairs owned by `rank` (remainder spread over the
    first ranks, like DistributedSampler without padding)."""
    if total < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad shard request total={total} rank={rank} world_size={world_size}")
    base, rem = divmod(total, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _reduce_device(world: World):
    return torch.device("cuda", world.local_rank) if world.backend == "nccl" else torch.device("cpu")


def barrier(world: World) -> None:
    if world.distributed:
        dist.barrier()


def max_over_ranks(world: World, value: float) -> float:
    """Max of a per-rank scalar (device time of the timed region): every multi-GPU number is the slowest rank's."""
    if not world.distributed:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_reduce_device(world))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(world: World, value: float) -> float:
    if not world.distributed:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_reduce_device(world))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def throughput(world: World, units_this_rank: float, elapsed_ms_this_rank: float) -> tuple[float, float]:
    """(whole-job units/s, ms) = all ranks' units / slowest rank's device time."""
    total_units = sum_over_ranks(world, units_this_rank)
    ms = max_over_ranks(world, elapsed_ms_this_rank)
    return (total_units / (ms * 1e-3) if ms > 0 else float("nan")), ms


def shutdown(world: World) -> None:
    if world.distributed and dist.is_initialized():
        dist.destroy_process_group()
