"""Multi-GPU plumbing: one process per GPU, each owning a contiguous path slice of every vector.

The path shards naturally: elementwise ops need no exchange; reductions and regression normal
equations all-reduce a handful of doubles inside the C runtime (ncclAllReduce on the compute stream). torch.distributed
is used only to hand the NCCL unique id to the other ranks and for barriers/timing in bench.py.
"""
from __future__ import annotations

import ctypes as C

from . import _capi as capi


def path_slice(numberOfPaths: int, rank: int, world_size: int, align: int = 4) -> tuple[int, int]:
    """Contiguous slice [p0, p1) of rank `rank`; boundaries are multiples of `align` (128-bit vector loads)."""
    per = -(-numberOfPaths // world_size)
    per = -(-per // align) * align
    p0 = min(rank * per, numberOfPaths)
    p1 = min(p0 + per, numberOfPaths)
    return p0, p1


def stream_word_offset(p0: int, numberOfTimeSteps: int, numberOfFactors: int) -> int:
    """First MT19937 word of path p0: the stream is consumed path-major, two 32-bit words per increment."""
    return 2 * numberOfTimeSteps * numberOfFactors * p0


def init_comm_from_torch() -> tuple[int, int]:
    """Create the runtime's NCCL communicator using an initialised torch.distributed process group for the id exchange."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    capi.ensure_init()
    if world == 1:
        return rank, world
    buf = C.create_string_buffer(capi.UNIQUE_ID_BYTES)
    if rank == 0:
        capi.check(capi.load().fmc_comm_get_unique_id(buf))
    obj = [bytes(buf.raw)]
    dist.broadcast_object_list(obj, src=0)
    capi.check(capi.load().fmc_comm_init(rank, world, obj[0]))
    return rank, world


def comm_destroy() -> None:
    if capi._lib is not None:
        capi.check(capi.load().fmc_comm_destroy())
