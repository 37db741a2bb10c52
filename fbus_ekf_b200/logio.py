"""Readers / writers of the reference's recorded-log text files, byte-compatible with what the C++
code writes under OPEN_DATA_RECORDING: one record per line, fields separated by one blank, every double through
`ofstream <<` at the default precision (i.e. printf "%g", 6 significant digits), marker ids as integers.

    imu.txt      t ax ay az gx gy gz                       raw (pre-IIR) samples           filter.cpp:28-34
    image.txt    t id px py pz qw qx qy qz                 marker pose in the left camera  vision.cpp:104-106
    corners.txt  t id + 16 stereo corner coordinates       water mode                      vision.cpp:111-119
                 t id + 4 x (X Y Z) triangulated corners   land mode (older writer)        vision.cpp:120-124
    fusion.txt   t p(3) q(wxyz) v(3) b_a(3) b_g(3)          filter output per frame         filter.cpp:241-246

Host-side I/O only; nothing here computes on the path."""
from __future__ import annotations

import os

import numpy as np

COLUMNS = {"imu": (7,), "image": (9,), "corners": (18, 14), "fusion": (17,)}
_ID_COLUMN = {"imu": None, "image": 1, "corners": 1, "fusion": None}


def read_log(path: str, kind: str) -> np.ndarray:
    """-> float64 [n, columns]; raises ValueError when the column count is not one the reference writes for `kind`."""
    if kind not in COLUMNS:
        raise ValueError(f"unknown log kind {kind!r} (one of {sorted(COLUMNS)})")
    if os.path.getsize(path) == 0:
        return np.zeros((0, COLUMNS[kind][0]))
    rows = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if rows.size == 0:
        return np.zeros((0, COLUMNS[kind][0]))
    if rows.shape[1] not in COLUMNS[kind]:
        raise ValueError(f"{path}: {rows.shape[1]} columns, a {kind} log has {' or '.join(map(str, COLUMNS[kind]))}")
    return rows


def read_imu_log(path):
    return read_log(path, "imu")


def read_image_log(path):
    return read_log(path, "image")


def read_corners_log(path):
    return read_log(path, "corners")


def read_fusion_log(path):
    return read_log(path, "fusion")


def format_row(row, id_column=None) -> str:
    """one record as `ofstream << a << " " << b ...` prints it"""
    return " ".join(("%d" % int(v)) if i == id_column else ("%g" % float(v)) for i, v in enumerate(row))


def write_log(path: str, rows: np.ndarray, kind: str, newline: str = "\n", append: bool = False) -> None:
    """Writes `rows` in the reference's text format.  `newline="\\r\\n"` reproduces the bundled (Windows-checkout) files
    byte for byte; `append=True` mirrors the reference's `std::ios::app` recording."""
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    if rows.size and rows.shape[1] not in COLUMNS[kind]:
        raise ValueError(f"a {kind} log has {' or '.join(map(str, COLUMNS[kind]))} columns, got {rows.shape[1]}")
    idc = _ID_COLUMN[kind]
    with open(path, "a" if append else "w", newline="") as fh:
        for r in rows:
            fh.write(format_row(r, idc) + newline)


def write_fusion_log(path, rows, newline="\n", append=False):
    write_log(path, rows, "fusion", newline, append)


def read_dataset(directory: str) -> dict:
    """every log present in a dataset directory (matlab/dataset/*/dataset-NN layout) -> {kind: rows}"""
    out = {}
    for kind in COLUMNS:
        p = os.path.join(directory, kind + ".txt")
        if os.path.exists(p):
            out[kind] = read_log(p, kind)
    return out
