"""Multi-GPU plumbing of the path: correspondence sets are independent, so a global batch is
split into contiguous per-rank shards with NO data-path collective; torch.distributed is used only to agree
on the timing (max over ranks) and to gather per-rank counters / checksums.  Backend-agnostic: NCCL on the
GPU box, gloo in the CPU tests (tests/test_shard_gloo.py)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `total` sets over `world` ranks; the first total % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad shard request: total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """The job's time is the slowest rank's time."""
    if _world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_counters(counters: Dict[str, float]) -> List[Dict[str, float]]:
    """Every rank's counter dict, in rank order (timing counters, output checksums)."""
    if _world() == 1:
        return [dict(counters)]
    out: List[Dict[str, float]] = [None] * _world()  # type: ignore[list-item]
    dist.all_gather_object(out, dict(counters))
    return out


def output_checksum(final_trans: torch.Tensor, final_labels: torch.Tensor) -> Dict[str, float]:
    """Order-independent digest of a shard's outputs, comparable across different shardings of one batch."""
    return {"sets": float(final_trans.shape[0]),
            "trans_sum": float(final_trans.double().sum()),
            "trans_abs": float(final_trans.double().abs().sum()),
            "inliers": float(final_labels.double().sum())}
