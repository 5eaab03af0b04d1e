"""One-process-per-GPU plumbing for the distributed solve.

The reference driver bootstraps NCCL over MPI (cuda/acg-cuda.c:1104-1122:
rank 0 calls ncclGetUniqueId, MPI_Bcast, ncclCommInitRank) and lets rank 0
partition the matrix and scatter the parts (:1516-1782).  This image has no
MPI; the same steps run over ``torch.distributed`` (gloo for the CPU-side
object broadcast, NCCL for the data path inside libacgb200):

* ``init_process()``     -- rendezvous from the torchrun environment
* ``nccl_comm()``        -- unique id from rank 0, broadcast, ``Comm.init_nccl``
* ``local_part()``       -- every rank builds the (synthetic) matrix, partitions
                            it with the same row->part map and keeps its part
* ``block_partition()``  -- geometric px*py*pz row->part map for stencil grids;
                            ``rowparts="metis"`` uses acgsymcsrmatrix_partition_rows
* ``local_stencil_part()`` / ``local_part_from_file()`` -- a rank's part built
                            directly (generator / streamed binary Matrix Market
                            file) without anybody holding the global matrix
"""
from __future__ import annotations

import os

import numpy as np


def init_process(backend: str | None = None):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, world, local


def nccl_comm(rank: int, world: int):
    """Communicator for libacgb200: a null comm for one process, else NCCL."""
    from .api import Comm
    if world == 1:
        return Comm()
    import torch.distributed as dist
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return Comm.init_nccl(world, rank, box[0])


def block_partition(nx: int, ny: int, nz: int, px: int, py: int, pz: int) -> np.ndarray:
    i = np.arange(nx * ny * nz, dtype=np.int64)
    x, y, z = i % nx, (i // nx) % ny, i // (nx * ny)
    return ((x * px // nx) + px * ((y * py // ny) + py * (z * pz // nz))).astype(np.int32)


def grid_factors(nparts: int) -> tuple[int, int, int]:
    """px*py*pz = nparts, as cubic as possible, larger factors on the slow axes."""
    best = (1, 1, nparts)
    for a in range(1, nparts + 1):
        if nparts % a:
            continue
        for b in range(a, nparts // a + 1):
            if (nparts // a) % b:
                continue
            c = nparts // a // b
            if c >= b and (c - a) < (best[2] - best[0]):
                best = (a, b, c)
    return best


def local_stencil_part(kind: int, nx: int, ny: int, nz: int, rank: int, world: int, eps: float = 0.0):
    """This rank's part of a 7/27-point stencil matrix on an nx*ny*nz box under the
    geometric block partition, generated directly (acgb200_stencil_part): no process
    ever holds the global matrix.  Array-for-array identical to
    ``local_part(..., block_partition(...))`` (tests/test_host_structs.py)."""
    from .api import SymCsrMatrix
    px, py, pz = grid_factors(world)
    return SymCsrMatrix.stencil_part(kind, nx, ny, nz, px, py, pz, rank).dsymv_init(eps)


def local_part(n, rows, cols, vals, rowparts, rank: int, world: int, eps: float = 0.0):
    """This rank's submatrix (full storage initialised), as the reference's
    rank 0 would have scattered it (acg/symcsrmatrix.c:685 + :1238)."""
    from .api import SymCsrMatrix
    A = SymCsrMatrix.init_real_double(n, rows, cols, vals)
    if world == 1:
        return A.dsymv_init(eps)
    if isinstance(rowparts, str) and rowparts == "metis":
        # METIS is deterministic for a fixed seed: every rank computes the same map
        rowparts, _ = A.partition_rows(world, kway=False, seed=0)
    parts = A.partition(world, rowparts)
    A.free()
    mine = parts[rank]
    for p, m in enumerate(parts):
        if p != rank:
            m.free()
    return mine.dsymv_init(eps)


def contiguous_partition(n: int, nparts: int) -> np.ndarray:
    """Row -> part map of nparts contiguous, equally sized row blocks."""
    return (np.arange(n, dtype=np.int64) * nparts // max(n, 1)).astype(np.int32)


def local_part_from_file(path: str, rowparts, rank: int, world: int, eps: float = 0.0):
    """This rank's part of the matrix in a binary Matrix Market file
    (acgb200_mtx_read_part): every rank streams the file and keeps the entries of
    its own rows -- the reference reads on rank 0 and scatters
    (cuda/acg-cuda.c:1297-1304, :1516-1782).  ``rowparts``: an array, or "rows"
    for contiguous row blocks."""
    from .api import SymCsrMatrix, mtx_info
    if world == 1:
        return SymCsrMatrix.read_mtx(path, binary=True).dsymv_init(eps)
    if isinstance(rowparts, str) and rowparts == "rows":
        rowparts = contiguous_partition(mtx_info(path)["nrows"], world)
    return SymCsrMatrix.read_mtx_part(path, world, rowparts, rank).dsymv_init(eps)


def comm_matrix(A, rank: int, world: int) -> np.ndarray:
    """world x world matrix of halo send counts, assembled from every rank's row
    (the driver's --output-comm-matrix, cuda/acg-cuda.c:1713-1775)."""
    row = A.comm_matrix_row(world)
    if world == 1:
        return row.reshape(1, 1)
    import torch.distributed as dist
    rows = [None] * world
    dist.all_gather_object(rows, row)
    return np.stack(rows)


def balanced_rows_partition(A, nparts: int) -> np.ndarray:
    """Contiguous row blocks holding about the same number of nonzeros of the full
    matrix each (the SpMV work), for inputs whose row lengths vary wildly: on an
    R-MAT graph equal row counts put most of the nonzeros on the first rank."""
    n = A.c.nprows
    rp = A.rowptr
    cols = A.colidx - A.c.rowidxbase
    deg = np.diff(rp).astype(np.int64)                       # packed entries of the row itself
    deg += np.bincount(cols, minlength=n)[:n]                # mirrored entries
    rows = np.repeat(np.arange(n), np.diff(rp))
    deg -= np.bincount(rows[cols == rows], minlength=n)[:n]  # the diagonal entry was counted twice
    cum = np.cumsum(deg)
    bounds = np.searchsorted(cum, cum[-1] * np.arange(1, nparts) / nparts, side="left")
    parts = np.zeros(n, np.int32)
    for b in bounds:
        parts[b + 1:] += 1
    return parts
