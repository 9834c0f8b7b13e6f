"""Host-side partitioning of the denoise workload over ranks (one process per GPU).

The reference is single-GPU; both partitionings below are net-new:
  * independent frame sequences, one context + stream per GPU, no data-path collective
    (BASELINE.json configs[4]);
  * one large frame split into contiguous row bands (BASELINE.json configs[3]): rows are contiguous in
    memory (pitch = W texels, reference include/extended_math.h:66-68), so a band plus its halo is one
    contiguous block.
torch.distributed is only the plumbing (rendezvous, max-over-ranks timing).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

# vertical reach of the passes below the temporal one (DESIGN.md "Row bands"):
#   variance 7x7 -> 3 rows; a-trous level l -> 2 * 2^l rows (+1 row for the 3x3 variance prefilter)
ATROUS_LEVEL_HALO = [2, 4, 8, 16, 32]
VARIANCE_HALO = 3


def assign_sequences(num_sequences: int, world_size: int):
    """sequence j -> rank j mod world_size; returns one list of sequence ids per rank."""
    if world_size < 1 or num_sequences < 0:
        raise ValueError("world_size >= 1 and num_sequences >= 0 required")
    return [list(range(r, num_sequences, world_size)) for r in range(world_size)]


@dataclass(frozen=True)
class Band:
    rank: int
    row0: int      # first owned row
    rows: int      # owned rows
    halo_top: int  # rows needed above row0 for a halo-recompute frame (clipped at the image border)
    halo_bot: int


def frame_halo(levels: int) -> int:
    """Rows of INPUT a band needs beyond its own rows so that every pass of a frame can be evaluated
    without mid-frame exchange (sum of the per-pass reaches; 62 + 1 + 3 = 66 for 5 levels)."""
    if not 0 <= levels <= len(ATROUS_LEVEL_HALO):
        raise ValueError("levels out of range")
    return sum(ATROUS_LEVEL_HALO[:levels]) + (1 if levels else 0) + VARIANCE_HALO


def row_bands(height: int, world_size: int, levels: int = 5):
    """Contiguous, near-equal row bands covering [0, height) exactly once."""
    if world_size < 1 or height < world_size:
        raise ValueError("need at least one row per rank")
    halo = frame_halo(levels)
    base, extra = divmod(height, world_size)
    bands, row0 = [], 0
    for r in range(world_size):
        rows = base + (1 if r < extra else 0)
        bands.append(Band(r, row0, rows, min(halo, row0), min(halo, height - (row0 + rows))))
        row0 += rows
    return bands


def max_over_ranks(value: float, device=None) -> float:
    """Multi-GPU timings are reported as the max over ranks (never wall clock of one rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_mpixels_per_s(pixels_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job throughput: all ranks' pixels over the slowest rank's time."""
    return sum_over_ranks(pixels_this_rank, device) / max_over_ranks(seconds_this_rank, device) / 1e6
