"""Selection surface of the interval join, mirrored from the reference's session layer
(sequila/sequila-core/src/session_context.rs, "SC"):

* :class:`Algorithm`      SC:62-72 enum + ``Cuda``; ``from_str`` = ``FromStr`` SC:85-104 (case-insensitive,
  same error text), ``str()`` = ``Display`` SC:106-120 — the text ``EXPLAIN`` prints as ``alg=...``
  (interval_join.rs:356-365).
* :class:`SequilaConfig`  SC:50-60 ``extensions_options!`` block, prefix ``sequila`` (SC:58-60).
* :func:`apply_set`       what DataFusion does with ``SET sequila.<key> TO <value>`` (ConfigField::set,
  SC:127-131).

Only ``Algorithm.Cuda`` executes in this repository: the CPU algorithms are the reference's own code and
are out of scope; selecting them here raises at execute, never silently runs on the GPU.
"""
from __future__ import annotations

import enum
import re
from dataclasses import dataclass


class ParseAlgorithmError(ValueError):
    """SC:74-83"""


class Algorithm(enum.Enum):
    Coitrees = "Coitrees"
    IntervalTree = "IntervalTree"
    ArrayIntervalTree = "ArrayIntervalTree"
    Lapper = "Lapper"
    SuperIntervals = "SuperIntervals"
    CoitreesNearest = "CoitreesNearest"
    CoitreesCountOverlaps = "CoitreesCountOverlaps"
    Cuda = "Cuda"  # the arm this repository adds next to coitrees and lapper
    CudaNearest = "CudaNearest"  # CoitreesNearest's semantics (one row per probe row) on the same GPU index

    @classmethod
    def default(cls) -> "Algorithm":
        return cls.Coitrees  # SC:64 #[default]

    @classmethod
    def from_str(cls, s: str) -> "Algorithm":
        table = {a.value.lower(): a for a in cls}
        try:
            return table[s.lower()]
        except KeyError:
            raise ParseAlgorithmError(f"Can't parse '{s}' as Algorithm") from None  # SC:98-101

    def __str__(self) -> str:
        return self.value


from dataclasses import field  # noqa: E402

# `sequila.cuda_*` keys = the tuning options of the `cuda` algorithm (include/sequila_cuda.h, sq_ctx_set_option); values are
# validated by the library when the exec node applies them, as ConfigField::set does for the reference's own fields
CUDA_KEYS = ("cuda_probe_layout", "cuda_staged_probe", "cuda_probe_block", "cuda_probe_tiles", "cuda_lookback_backoff_ns", "cuda_rows_per_bin",
             "cuda_right_idx_wire", "cuda_l2_persist_mb", "cuda_scan_dict_capacity", "cuda_exec_trace", "cuda_pipeline_depth",
             "cuda_coalesce_rows", "cuda_rank_count", "cuda_build_ids", "cuda_build_sort")


@dataclass
class SequilaConfig:
    """SC:50-56.  Keys are addressed as ``sequila.<field>``."""
    prefer_interval_join: bool = True
    interval_join_algorithm: Algorithm = Algorithm.Coitrees
    interval_join_low_memory: bool = False
    cuda: dict = field(default_factory=dict)  # sequila.cuda_* keys that were SET, applied to the exec node's context

    PREFIX = "sequila"

    def set(self, key: str, value: str) -> None:
        key = key.strip().lower()
        if key.startswith(self.PREFIX + "."):
            key = key[len(self.PREFIX) + 1:]
        if key == "interval_join_algorithm":
            self.interval_join_algorithm = Algorithm.from_str(value)
        elif key in ("prefer_interval_join", "interval_join_low_memory"):
            v = value.strip().lower()
            if v not in ("true", "false"):
                # DataFusion's bool ConfigField: str::parse::<bool>() failure
                raise ValueError(f"Error parsing {value} as bool")
            setattr(self, key, v == "true")
        elif key in CUDA_KEYS:
            self.cuda[key] = value.strip()
        else:
            raise KeyError(f'Config value "{key}" not found on SequilaConfig')


_SET_RE = re.compile(r"^\s*SET\s+([A-Za-z_][\w.]*)\s*(?:TO|=)\s*(.+?)\s*;?\s*$", re.I)


def apply_set(config: SequilaConfig, statement: str) -> None:
    """``SET sequila.interval_join_algorithm TO cuda`` (README.md:30-41 of the reference)."""
    m = _SET_RE.match(statement)
    if not m:
        raise ValueError(f"not a SET statement: {statement!r}")
    key, value = m.group(1), m.group(2)
    if len(value) >= 2 and value[0] == value[-1] and value[0] in "'\"":
        value = value[1:-1]
    if not key.lower().startswith(SequilaConfig.PREFIX + "."):
        raise KeyError(f"unknown configuration namespace in {key!r}")
    config.set(key, value)
