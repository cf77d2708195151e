"""Seeded synthetic relations for the BASELINE.json configs, generated on the device by K6 (hjGenerate).

The reference's generators (shared_stuff/shared.cpp:35-116) are unseeded rand(); these are counter-based and
integer-only, and bit-identical to the oracle's (oracle/oracle_join.c gen_one), so CPU and GPU see the same inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import torch

from . import _lib

KIND_INDEX, KIND_UNIQUE, KIND_UNIFORM, KIND_MIXED, KIND_FK, KIND_ZIPF = range(6)
ODD_MUL64 = 0x9E3779B97F4A7C15          # odd -> multiplication is a bijection mod 2^64 (spreads keys over int64)


@dataclass(frozen=True)
class RelationSpec:
    n: int
    key_bytes: int
    kind: int
    seed: int
    lo: int = 0
    domain: int = 1
    p16: int = 0
    key_mul: int = 0

    @property
    def dtype(self) -> torch.dtype:
        return torch.int32 if self.key_bytes == 4 else torch.int64


@dataclass(frozen=True)
class JoinConfig:
    name: str
    build: RelationSpec
    probe: RelationSpec
    expected_out: int | None        # analytic result size when the generators pin it
    note: str = ""


def generate(spec: RelationSpec, device="cuda", index_base: int = 0, n_local: int | None = None) -> torch.Tensor:
    """Rows [index_base, index_base + n_local) of the relation ``spec`` describes (all of it by default)."""
    n = spec.n - index_base if n_local is None else n_local
    out = torch.empty(n, dtype=spec.dtype, device=device)
    rc = _lib.load().hjGenerate(out.data_ptr(), n, spec.key_bytes, spec.kind, spec.seed, spec.lo, spec.domain, spec.p16,
                                spec.key_mul, index_base, spec.n, torch.cuda.current_stream().cuda_stream)
    _lib.check_status(rc, "hjGenerate")
    return out


def generate_at(spec: RelationSpec, row_ids: torch.Tensor) -> torch.Tensor:
    """Keys of the rows ``row_ids`` (int32 patterns of u32 row ids) of the relation ``spec`` describes (hjGenerateAt)."""
    out = torch.empty(row_ids.numel(), dtype=spec.dtype, device=row_ids.device)
    rc = _lib.load().hjGenerateAt(out.data_ptr(), row_ids.data_ptr(), row_ids.numel(), spec.key_bytes, spec.kind, spec.seed, spec.lo, spec.domain, spec.p16,
                                  spec.key_mul, spec.n, torch.cuda.current_stream().cuda_stream)
    _lib.check_status(rc, "hjGenerateAt")
    return out


def config(name: str, scale_log2: int = 0) -> JoinConfig:
    """The five BASELINE.json configs. ``scale_log2`` < 0 shrinks both relations by 2^k."""
    s = scale_log2

    def sh(x):
        return max(1, x >> -s) if s < 0 else x << s
    if name == "C1":       # 1K x 4K i32, unique build keys, ~50 % hits
        return JoinConfig("C1", RelationSpec(1024, 4, KIND_UNIQUE, 1, 0, 1024), RelationSpec(4096, 4, KIND_UNIFORM, 2, 0, 2048), None,
                          "join_v1.mlir sizes are compile-time constants; 1Kx4K is check()-sized")
    if name == "C2":       # 16M x 256M i32, unique build, 100 % match
        nR, nS = sh(1 << 24), sh(1 << 28)
        return JoinConfig("C2", RelationSpec(nR, 4, KIND_UNIQUE, 42, 0, nR), RelationSpec(nS, 4, KIND_UNIFORM, 43, 0, nR), nS,
                          "HBM-resident table, every probe key hits exactly one build row")
    if name == "C2S":      # config 2 with sparse keys: the same permutation spread over int32 by an odd multiplier -> no dense range to ride
        nR, nS = sh(1 << 24), sh(1 << 28)
        return JoinConfig("C2S", RelationSpec(nR, 4, KIND_UNIQUE, 42, 0, nR, 0, 0x9E3779B1), RelationSpec(nS, 4, KIND_UNIFORM, 43, 0, nR, 0, 0x9E3779B1), nS,
                          "config 2's sizes and match pattern, keys multiplied by an odd constant mod 2^32 (a bijection): only the hash-table paths apply")
    if name == "REF10M":   # join-performances.md:3-6 / :16-19
        n = sh(10_000_000)
        return JoinConfig("REF10M", RelationSpec(n, 4, KIND_UNIFORM, 50, 1, 100_000), RelationSpec(n, 4, KIND_UNIFORM, 51, 1, 100_000), None,
                          "the reference's published shape 1: 10M x 10M rows, keys uniform in [1, 100k], ~1e9 result pairs")
    if name == "REF100M":  # join-performances.md:8-11 / :21-24
        n = sh(100_000_000)
        return JoinConfig("REF100M", RelationSpec(n, 4, KIND_UNIFORM, 52, 1, 1_000_000_000), RelationSpec(n, 4, KIND_UNIFORM, 53, 1, 1_000_000_000), None,
                          "the reference's published shape 2: 100M x 100M rows, keys uniform in [1, 1e9], ~1e7 result pairs")
    if name == "C3":       # 1M x 1B i32, 10 % selectivity, L2-resident table
        nR, nS = sh(1 << 20), sh(1 << 30)
        return JoinConfig("C3", RelationSpec(nR, 4, KIND_UNIQUE, 44, 0, nR), RelationSpec(nS, 4, KIND_MIXED, 45, 0, nR, 6554), None,
                          "10 % of probe keys from the build key set, 90 % from a disjoint range")
    if name == "C4":       # i64 FK 1:4 build, Zipf(1.0) probe
        D, nR, nS = sh(1 << 23), sh(1 << 25), sh(1 << 27)
        return JoinConfig("C4", RelationSpec(nR, 8, KIND_FK, 46, 0, D, 0, ODD_MUL64), RelationSpec(nS, 8, KIND_ZIPF, 47, 0, D, 0, ODD_MUL64),
                          nS * (nR // D), "every probe key matches nR/D build rows; probe keys Zipf(1.0) over D")
    if name == "C5":       # 2e9 x 2e9 i64, unique x unique
        n = max(1, int(2_000_000_000 * 2.0 ** s))
        return JoinConfig("C5", RelationSpec(n, 8, KIND_UNIQUE, 48, 0, n, 0, ODD_MUL64), RelationSpec(n, 8, KIND_UNIQUE, 49, 0, n, 0, ODD_MUL64), n,
                          "radix-partitioned, every key on both sides exactly once")
    raise KeyError(name)


def shrink(cfg: JoinConfig, n_build: int, n_probe: int) -> JoinConfig:
    """Same distributions at explicit sizes (parity tests)."""
    b, p = cfg.build, cfg.probe
    dom_b = n_build if b.kind == KIND_UNIQUE else max(1, b.domain * n_build // b.n)
    dom_p = dom_b if p.domain == b.domain else max(1, p.domain * n_build // b.n)
    return JoinConfig(cfg.name, replace(b, n=n_build, domain=dom_b), replace(p, n=n_probe, domain=dom_p), None, cfg.note)
