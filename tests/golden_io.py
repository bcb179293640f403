"""Golden-vector plumbing: a hex-float dump format, its loader, and a per-stage comparator.

Reference-independent on purpose.  The day base-cpp is staged, an instrumented harness
(outside the reference tree, including oracle/b9_dump.h) calls the reference's own
likelihood on a fixed-seed cluster and writes one record per (stage, star) with the values
as C99 hex floats ("%a": every bit of the double survives a text round trip, on any libc).
This module reads those files, and compares a candidate implementation with them stage by
stage, reporting the FIRST divergence in pipeline order — the place to start debugging —
rather than a bare pass/fail on the final log-posterior.  Nothing here knows what the
stages are or what the numbers mean.

File format (text, line oriented, '#' starts a comment line):

    b9dump 1
    meta <key> <value...>                      any number, e.g. commit, seed, config, tolerance
    rec <stage> <star> <n>                     stage: [A-Za-z0-9_.-]+   star: integer >= -1
    <n hex floats, whitespace separated, any number of lines>
    ...
    end <record count>

`star` is -1 for a per-cluster record (the summed log-posterior, a proposal's parameters).
inf / -inf / nan are spelled as printf("%a") spells them.
"""
from __future__ import annotations

import dataclasses
import math
import re
from pathlib import Path
from typing import Iterable, Iterator

import numpy as np

MAGIC = "b9dump 1"
_STAGE = re.compile(r"^[A-Za-z0-9_.\-]+$")


@dataclasses.dataclass(frozen=True)
class Record:
    stage: str
    star: int
    values: np.ndarray            # float64, 1-D


@dataclasses.dataclass
class Dump:
    meta: dict[str, str]
    records: list[Record]

    def stages(self) -> list[str]:
        """Stage names in first-appearance order: the pipeline order the harness wrote them in."""
        seen: dict[str, None] = {}
        for r in self.records:
            seen.setdefault(r.stage)
        return list(seen)

    def by_key(self) -> dict[tuple[str, int], Record]:
        out: dict[tuple[str, int], Record] = {}
        for r in self.records:
            if (r.stage, r.star) in out:
                raise ValueError(f"duplicate record {r.stage!r} star {r.star}")
            out[(r.stage, r.star)] = r
        return out


def _hex(x: float) -> str:
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return float(x).hex()


def _unhex(tok: str) -> float:
    t = tok.lower()
    if t in ("nan", "-nan", "+nan"):
        return math.nan
    if t in ("inf", "+inf", "infinity"):
        return math.inf
    if t in ("-inf", "-infinity"):
        return -math.inf
    return float.fromhex(tok)     # accepts glibc's "0x1.8p+1" and Python's "0x1.8000000000000p+1"


def dump(path, records: Iterable[Record], meta: dict[str, str] | None = None, per_line: int = 4) -> None:
    lines = [MAGIC]
    for k, v in (meta or {}).items():
        if not _STAGE.match(k) or "\n" in str(v):
            raise ValueError(f"bad meta entry {k!r}")
        lines.append(f"meta {k} {v}")
    n = 0
    for r in records:
        if not _STAGE.match(r.stage) or r.star < -1:
            raise ValueError(f"bad record key {r.stage!r} {r.star}")
        v = np.asarray(r.values, dtype=np.float64).ravel()
        lines.append(f"rec {r.stage} {r.star} {v.size}")
        for i in range(0, v.size, per_line):
            lines.append(" ".join(_hex(x) for x in v[i:i + per_line]))
        n += 1
    lines.append(f"end {n}")
    Path(path).write_text("\n".join(lines) + "\n")


def _tokens(lines: Iterator[tuple[int, str]], want: int, where: str) -> list[float]:
    vals: list[float] = []
    while len(vals) < want:
        try:
            no, line = next(lines)
        except StopIteration:
            raise ValueError(f"{where}: file ends inside a record ({len(vals)}/{want} values)") from None
        if line.startswith(("rec ", "end ", "meta ")):
            raise ValueError(f"{where}: line {no}: record is short ({len(vals)}/{want} values)")
        try:
            vals.extend(_unhex(t) for t in line.split())
        except ValueError:
            raise ValueError(f"{where}: line {no}: not a hex float: {line!r}") from None
    if len(vals) != want:
        raise ValueError(f"{where}: record is long ({len(vals)}/{want} values)")
    return vals


def load(path) -> Dump:
    where = str(path)
    raw = Path(path).read_text().splitlines()
    lines = ((i + 1, l.strip()) for i, l in enumerate(raw) if l.strip() and not l.lstrip().startswith("#"))
    try:
        _, first = next(lines)
    except StopIteration:
        raise ValueError(f"{where}: empty file") from None
    if first != MAGIC:
        raise ValueError(f"{where}: not a b9dump version-1 file (first line {first!r})")
    meta: dict[str, str] = {}
    records: list[Record] = []
    ended = False
    for no, line in lines:
        if ended:
            raise ValueError(f"{where}: line {no}: content after 'end'")
        head, _, rest = line.partition(" ")
        if head == "meta":
            k, _, v = rest.partition(" ")
            meta[k] = v
        elif head == "rec":
            parts = rest.split()
            if len(parts) != 3 or not _STAGE.match(parts[0]):
                raise ValueError(f"{where}: line {no}: bad record header {line!r}")
            stage, star, n = parts[0], int(parts[1]), int(parts[2])
            if star < -1 or n < 0:
                raise ValueError(f"{where}: line {no}: bad star/count in {line!r}")
            vals = _tokens(lines, n, where)
            records.append(Record(stage, star, np.array(vals, dtype=np.float64)))
        elif head == "end":
            if int(rest) != len(records):
                raise ValueError(f"{where}: 'end {rest}' but {len(records)} records read (truncated?)")
            ended = True
        else:
            raise ValueError(f"{where}: line {no}: unknown line {line!r}")
    if not ended:
        raise ValueError(f"{where}: no 'end' line (truncated?)")
    return Dump(meta, records)


# ------------------------------------------------------------------ comparison
def mixed_err(got: np.ndarray, want: np.ndarray) -> np.ndarray:
    """|got-want| / max(1, |want|), elementwise; 0 where both are the same inf or both nan."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    same_special = (got == want) | (np.isnan(got) & np.isnan(want))
    with np.errstate(invalid="ignore"):
        err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    err = np.where(same_special, 0.0, err)
    return np.where(np.isnan(err), np.inf, err)       # inf vs finite, nan vs number: infinitely wrong


def ulp_distance(a, b) -> np.ndarray:
    ia = np.asarray(a, dtype=np.float64).view(np.int64).copy()
    ib = np.asarray(b, dtype=np.float64).view(np.int64).copy()
    ia[ia < 0] = np.int64(-(2 ** 63)) - ia[ia < 0]
    ib[ib < 0] = np.int64(-(2 ** 63)) - ib[ib < 0]
    return np.abs(ia - ib)


@dataclasses.dataclass(frozen=True)
class Divergence:
    stage: str
    star: int
    index: int
    got: float
    want: float
    err: float
    ulps: int

    def __str__(self) -> str:
        return (f"stage {self.stage!r} star {self.star} element {self.index}: got {self.got!r} "
                f"({float(self.got).hex() if math.isfinite(self.got) else self.got}) want {self.want!r} "
                f"({float(self.want).hex() if math.isfinite(self.want) else self.want}) "
                f"err {self.err:.3e} of max(1,|want|), {self.ulps} ulp")


@dataclasses.dataclass
class Report:
    ok: bool
    first: Divergence | None                  # first failing element in pipeline (stage) order
    per_stage: dict[str, dict]                # stage -> {records, elements, max_err, max_ulps, failures}
    missing: list[tuple[str, int]]            # keys in `want` absent from `got`
    extra: list[tuple[str, int]]              # keys in `got` absent from `want`
    shape_mismatch: list[tuple[str, int, int, int]]

    def summary(self) -> str:
        lines = []
        for st, s in self.per_stage.items():
            lines.append(f"{st:24s} records {s['records']:6d} elements {s['elements']:9d} "
                         f"max err {s['max_err']:.3e} max ulp {s['max_ulps']:d} failures {s['failures']}")
        if self.missing:
            lines.append(f"missing {len(self.missing)} record(s), first {self.missing[0]}")
        if self.extra:
            lines.append(f"unexpected {len(self.extra)} record(s), first {self.extra[0]}")
        if self.shape_mismatch:
            lines.append(f"{len(self.shape_mismatch)} record(s) with the wrong length, first {self.shape_mismatch[0]}")
        lines.append("OK" if self.ok else f"FIRST DIVERGENCE: {self.first}" if self.first else "FAILED (structure)")
        return "\n".join(lines)


def compare(got: Dump, want: Dump, tol: float | dict[str, float] = 1e-10, bit_exact: Iterable[str] = ()) -> Report:
    """Stage by stage in `want`'s pipeline order.  `tol` bounds mixed_err (a dict gives per-stage
    bounds, key "*" the default); stages named in `bit_exact` must match in every bit."""
    g, w = got.by_key(), want.by_key()
    exact = set(bit_exact)
    tol_of = (lambda st: tol.get(st, tol.get("*", 1e-10))) if isinstance(tol, dict) else (lambda st: tol)
    missing = [k for k in w if k not in g]
    extra = [k for k in g if k not in w]
    shape, first, per_stage = [], None, {}
    for st in want.stages():
        stat = {"records": 0, "elements": 0, "max_err": 0.0, "max_ulps": 0, "failures": 0}
        for (s, star), rec in w.items():
            if s != st or (s, star) not in g:
                continue
            a, b = g[(s, star)].values, rec.values
            if a.size != b.size:
                shape.append((s, star, a.size, b.size))
                continue
            stat["records"] += 1
            stat["elements"] += b.size
            if b.size == 0:
                continue
            err, ulps = mixed_err(a, b), ulp_distance(a, b)
            nan_pair = np.isnan(a) & np.isnan(b)
            ulps = np.where(nan_pair, 0, ulps)
            bad = (ulps != 0) if st in exact else (err > tol_of(st))
            finite = err[np.isfinite(err)]
            stat["max_err"] = max(stat["max_err"], float(finite.max()) if finite.size else 0.0,
                                  math.inf if np.isinf(err).any() else 0.0)
            stat["max_ulps"] = max(stat["max_ulps"], int(ulps.max()))
            stat["failures"] += int(bad.sum())
            if bad.any() and first is None:
                i = int(np.flatnonzero(bad)[0])
                first = Divergence(s, star, i, float(a[i]), float(b[i]), float(err[i]), int(ulps[i]))
        per_stage[st] = stat
    ok = first is None and not missing and not extra and not shape
    return Report(ok, first, per_stage, missing, extra, shape)
