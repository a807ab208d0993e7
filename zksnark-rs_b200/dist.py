"""Multi-GPU plumbing for a sharded proof: one process per GPU, torch.distributed for the exchange.

The MSM base vectors (sigma_g1.xi / xi_t / sum_delta, sigma_g2.xi) are sharded by contiguous point
ranges; each rank produces its partial sums of A, B, C (32 limbs, `zkb_prove_partial`), the records are
all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests) and folded by `zkb_prove_combine`.
Elliptic-curve addition is not an NCCL reduction operator, so the "allreduce" of the north star is
all-gather + on-device fold; 256 bytes per rank cross the fabric per proof.
"""

from __future__ import annotations

import numpy as np

PARTIAL_LIMBS = 32


def shard_range(length: int, rank: int, world: int):
    """[lo, hi) of a length-`length` vector owned by `rank` -- mirrors shard() in csrc/crs.cu."""
    return length * rank // world, length * (rank + 1) // world


def all_gather_partials(part: np.ndarray, device=None) -> np.ndarray:
    """part: (32,) uint64 from prove_partial -> (world, 32) uint64, identical on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    part = np.ascontiguousarray(part, dtype=np.uint64).reshape(PARTIAL_LIMBS)
    if world == 1:
        return part.reshape(1, PARTIAL_LIMBS).copy()
    src = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        src = src.to(device)
    out = torch.empty(PARTIAL_LIMBS * world, dtype=torch.int64, device=src.device)
    dist.all_gather_into_tensor(out, src)
    return out.cpu().numpy().view(np.uint64).reshape(world, PARTIAL_LIMBS)
