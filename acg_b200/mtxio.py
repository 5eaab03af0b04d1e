"""Matrix Market output in the two encodings the reference driver reads
(acg/mtxfile.c): plain text and aCG's binary variant (``--binary``: the text
header and size line, then ``rowidx[nnz]``, ``colidx[nnz]`` as 1-based int32 and
``a[nnz]`` as float64, acg/mtxfile.c:1107-1127; tools mtx2bin/mtx2bin.c:538-549
produce the same layout).  Input is the upper-triangle COO of acg_b200.matgen."""
from __future__ import annotations

import numpy as np


def write_symmetric(path: str, n: int, rows, cols, vals, binary: bool = True) -> None:
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    with open(path, "wb") as f:
        f.write(b"%%MatrixMarket matrix coordinate real symmetric\n")
        f.write(f"{n} {n} {len(vals)}\n".encode())
        if binary:
            (rows + 1).tofile(f)
            (cols + 1).tofile(f)
            vals.tofile(f)
        else:
            for i, j, a in zip(rows, cols, vals):
                f.write(f"{i + 1} {j + 1} {a:.17g}\n".encode())


def write_vector(path: str, x, binary: bool = True) -> None:
    x = np.ascontiguousarray(x, np.float64)
    with open(path, "wb") as f:
        f.write(b"%%MatrixMarket vector array real general\n")
        f.write(f"{len(x)}\n".encode())
        if binary:
            x.tofile(f)
        else:
            for a in x:
                f.write(f"{a:.17g}\n".encode())


def write_comm_matrix(path: str, counts) -> None:
    """The communication matrix in the form the reference driver prints it
    (cuda/acg-cuda.c:1741-1748): "matrix coordinate integer general", part numbers
    1-based in the file (acg/mtxfile.c:1471-1472), one entry per (sender, recipient) pair."""
    counts = np.asarray(counts)
    p, q = np.nonzero(counts)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n")
        f.write(f"{counts.shape[0]} {counts.shape[1]} {len(p)}\n")
        for i, j in zip(p, q):
            f.write(f"{i + 1} {j + 1} {int(counts[i, j])}\n")


def write_rowparts(path: str, rowparts) -> None:
    """Row -> part map as the reference's mtxpartition writes it and its driver reads it
    with --partition=FILE (cuda/acg-cuda.c:1542-1640): "vector array integer general",
    one 1-based part number per row."""
    rowparts = np.asarray(rowparts)
    with open(path, "w") as f:
        f.write("%%MatrixMarket vector array integer general\n")
        f.write(f"{len(rowparts)}\n")
        f.write("\n".join(str(int(p) + 1) for p in rowparts))
        f.write("\n")


def read_rowparts(path: str) -> np.ndarray:
    """Inverse of write_rowparts; also accepts "matrix array integer general" with one column."""
    with open(path) as f:
        head = f.readline().split()
        if len(head) < 5 or head[0] != "%%MatrixMarket" or head[2] != "array" or head[1] not in ("vector", "matrix"):
            raise ValueError(f"{path}: expected a Matrix Market vector in array format")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        dims = [int(t) for t in line.split()]
        if head[1] == "matrix" and (len(dims) != 2 or dims[1] != 1):
            raise ValueError(f"{path}: expected one column")
        vals = np.array(f.read().split(), dtype=np.int64)
    if len(vals) != dims[0]:
        raise ValueError(f"{path}: {dims[0]} entries announced, {len(vals)} found")
    return (vals - 1).astype(np.int32)
