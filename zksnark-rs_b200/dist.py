"""Multi-GPU plumbing: one process per GPU.

ONE PROOF OVER ALL RANKS (the default layout of bench.py at N > 1): `connect` builds the in-library communicator
(zkb_comm_*: every rank's exchange window in HBM is mapped by its peers, and the kernels exchange the sharded
transforms' columns and the partial sums by direct stores over NVLink) -- torch.distributed only carries the
128-byte window handles once at start-up.  `layout_s_index` is the host-side mirror of the device's "strided block"
ownership (csrc/shard.cu).

Older host-driven helpers, kept for the stand-alone sweeps (tools/msm_multi_gpu.py, tools/ntt_multi_gpu.py):

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


def all_gather_limbs(arr: np.ndarray, device=None) -> np.ndarray:
    """(k, 4) uint64 limbs on every rank -> (world, k, 4), rank order (gloo on CPU tensors, NCCL on `device`)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    if world == 1:
        return a.reshape((1,) + a.shape).copy()
    src = torch.from_numpy(a.view(np.int64).reshape(-1).copy())
    if device is not None:
        src = src.to(device)
    out = torch.empty(src.numel() * world, dtype=torch.int64, device=src.device)
    dist.all_gather_into_tensor(out, src)
    return out.cpu().numpy().view(np.uint64).reshape((world,) + a.shape)


def ntt_shard_layout(log_n: int, rank: int, world: int):
    """Who holds what in the outer-dimension-sharded transform: (input = x[rank::world], output range [k0, k0 + count))."""
    log_g = world.bit_length() - 1
    if 1 << log_g != world or log_g > log_n:
        raise ValueError("ntt_sharded: world must be a power of two not larger than the transform")
    sub = (1 << log_n) >> log_g
    return slice(rank, None, world), rank * sub, sub


def ntt_sharded(ctx, x_sub: np.ndarray, log_n: int, inverse: bool = False) -> np.ndarray:
    """One size-2^log_n transform with its outer dimension sharded over the ranks (world a power of two).

    x_sub: this rank's decimated subsequence x[rank::world] as (n/world, 4) uint64 canonical limbs.  Returns this
    rank's slice X[rank*n/world : (rank+1)*n/world] (same layout).  Local size-n/world transform (zkb_ntt_fr), ONE
    all-gather of the partial transforms (n*32 bytes arrive at every rank; NCCL on GPUs), then the
    final coefficient reduction X[k] = sum_g w^(g k) Y_g[k mod n/world] on the device (zkb_ntt_combine).
    """
    import importlib

    import torch
    import torch.distributed as dist

    zg = importlib.import_module(__package__ + ".groth16")
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    log_g = world.bit_length() - 1
    _, k0, sub = ntt_shard_layout(log_n, rank, world)
    assert x_sub.shape == (sub, 4)
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = torch.from_numpy(np.ascontiguousarray(x_sub, dtype=np.uint64).view(np.int64)).to(dev)
    if log_n - log_g >= 1:
        zg.ntt_dev(ctx, mine.data_ptr(), log_n - log_g, inverse)
    parts = torch.empty((world, sub, 4), dtype=torch.int64, device=dev)
    if world > 1:
        torch.cuda.synchronize()
        dist.all_gather_into_tensor(parts.view(-1), mine.view(-1))
        torch.cuda.synchronize()
    else:
        parts[0] = mine
    out = torch.empty((sub, 4), dtype=torch.int64, device=dev)
    zg.ntt_combine(ctx, parts.data_ptr(), log_n, log_g, inverse, k0, sub, out.data_ptr())
    return out.cpu().numpy().view(np.uint64)


# ------------------------------------------------------------------------------------------------
# in-library communicator (csrc/shard.cu)
def layout_s_index(n: int, rank: int, world: int) -> np.ndarray:
    """Global coefficient index of every local index of layout S on `rank`: s = k1*q + t <-> (rank*q + t) + (n/world)*k1.
    This is how u_sum, v_sum, h come out of a sharded proof's polynomial stage and how zkb_setup_shard /
    zkb_crs_upload_shard shard sigma_g1.xi, sigma_g1.xi_t (minus index n-1) and sigma_g2.xi."""
    m = n // world
    q = m // world
    if q * world * world != n:
        raise ValueError("layout S needs n to be a multiple of world^2")
    s = np.arange(m, dtype=np.int64)
    k1, t = s // q, s % q
    return rank * q + t + m * k1


def layout_d_index(n: int, rank: int, world: int) -> np.ndarray:
    """Layout D (decimated): local index i <-> global index rank + world*i (the gates a rank evaluates)."""
    return rank + world * np.arange(n // world, dtype=np.int64)


def connect(ctx, max_log_n: int, device=None):
    """Create this rank's communicator end and connect it to all ranks of the default torch.distributed group
    (any backend: only `world` x 128 bytes of window handles travel, once)."""
    import importlib

    import torch
    import torch.distributed as dist

    zg = importlib.import_module(__package__ + ".groth16")
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    comm = zg.Comm.create(ctx, rank, world, max_log_n)
    if world == 1:
        return comm.connect(comm.handle)
    mine = torch.frombuffer(bytearray(comm.handle), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(world * zg.COMM_HANDLE_BYTES, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    comm.connect(bytes(out.cpu().numpy().tobytes()))
    dist.barrier()  # nobody starts writing into a window before every rank has mapped it
    return comm
