"""Mirror of the reference's `buzhash` package surface (github.com/pbs-plus/pxar/buzhash).

The only call the reference makes is ``buzhash.NewConfig(4096)`` at
internal/pxarmount/commit.go:302-305 and it ignores the error.  The integer is the
average chunk size in KiB in the upstream client's convention (4096 -> 4 MiB, which is
also BASELINE.json's "4 MiB avg chunk"); ``NewConfigBytes`` exists for the other
reading.
"""
from __future__ import annotations

import numpy as np

from ._lib import Cfg
from .engine import make_config

Config = Cfg


def NewConfig(avg_kib: int, table: np.ndarray | None = None) -> Config:
    """buzhash.NewConfig(avgKiB) -> (Config, error); here: raises on error."""
    if avg_kib <= 0 or avg_kib > (1 << 19):
        raise ValueError(f"buzhash: invalid average chunk size {avg_kib} KiB")
    return make_config(avg_kib << 10, table)


def NewConfigBytes(avg_bytes: int, table: np.ndarray | None = None) -> Config:
    return make_config(avg_bytes, table)
