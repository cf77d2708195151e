"""Generates tests/golden/check_verdicts.json by running the REFERENCE's own check() (shared_stuff/shared.cpp:129-172,
compiled unmodified into oracle/_ref/shared.so by oracle/Makefile) on the fixtures the reference makes derivable
(SURVEY.md section 8c). Run in the build container, where /root/reference exists:

    make -C oracle && python tests/golden/make_golden.py

Each case stores the inputs, a candidate result and the verdict the reference returned (1 / 0 / -1). The candidate
results are nested-loop joins computed here in numpy (row-major i, j order like shared.cpp:154-165), some of them
deliberately corrupted, so the verdicts pin all three outcomes."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.binding import load_reference_check  # noqa: E402


def nested(R, S):
    i, j = np.nonzero(np.asarray(R)[:, None] == np.asarray(S)[None, :])
    return i.astype(np.int32), j.astype(np.int32)


def main():
    ref_check = load_reference_check()
    if ref_check is None:
        raise SystemExit("oracle/_ref/shared.so missing: run `make -C oracle` where /root/reference exists")
    rng = np.random.default_rng(20261018)
    cases = []

    def add(name, R, S, outR, outS, why):
        R, S = np.asarray(R, np.int32), np.asarray(S, np.int32)
        outR, outS = np.asarray(outR, np.int32), np.asarray(outS, np.int32)
        v = ref_check(R, S, outR, outS)
        cases.append(dict(name=name, why=why, R=R.tolist(), S=S.tolist(), outR=outR.tolist(), outS=outS.tolist(), verdict=v))

    # KAT-dup: shared.cpp:75 (R all 10) x shared.cpp:103-114 (S fixture), sized for join_v2.ll's 12x12 snapshot
    R = [10] * 12
    S = [10, 10, 3, 2, 3, 3, 5, 3, 5, 2, 2, 3]
    a, b = nested(R, S)
    add("kat_dup", R, S, a, b, "24 pairs {(i,0),(i,1)}")
    add("kat_dup_shuffled", R, S, a[::-1], b[::-1], "order must not matter (shared.cpp:168-169)")
    add("kat_dup_truncated", R, S, a[:-1], b[:-1], "more matches than result rows -> -1 (shared.cpp:158-160)")
    bad = a.copy(); bad[3] = (bad[3] + 1) % 12
    add("kat_dup_wrong_row", R, S, bad, b, "one build row wrong -> 0")
    add("kat_dup_padded", R, S, np.concatenate([a, [0]]), np.concatenate([b, [0]]), "result longer than the join: reference pads its own side with (0,0)")
    add("kat_dup_padded_bad", R, S, np.concatenate([a, [5]]), np.concatenate([b, [7]]), "extra non-(0,0) row -> 0")
    # KAT-none: shared.cpp:75 x shared.cpp:99 -> empty result (join_v1.mlir:635-644)
    add("kat_none", [10] * 12, [1] * 12, [], [], "resultSize == 0 branch")
    add("kat_none_nonempty_result", [10] * 12, [1] * 12, [0], [0], "a (0,0) row against an empty join: equals the padding")
    # KAT-index: initRelationIndex on both sides (shared.cpp:35-41)
    R, S = np.arange(10), np.arange(10)
    a, b = nested(R, S)
    add("kat_index_10x10", R, S, a, b, "pairs (i,i); 10x10 is join_v1.ll's snapshot shape (join_v1.ll:12-14)")
    R, S = np.arange(7), np.arange(12)
    a, b = nested(R, S)
    add("kat_index_7x12", R, S, a, b, "out = min(|R|,|S|)")
    # nested-loop.mlir:7-24,208-212: two 20-row tables, key column val = i + j with j the column -> keys i and i
    R, S = np.arange(20), np.arange(20)
    a, b = nested(R, S)
    add("kat_nested_20x20", R, S, a, b, "20 matches (i,i)")
    # empty relations
    add("empty_build", [], [1, 2, 3], [], [], "zero-length build memref")
    add("empty_probe", [1, 2, 3], [], [], [], "zero-length probe memref")
    # seeded random cases, duplicates on both sides, negative keys and extreme values
    for k in range(6):
        nR, nS = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        R = rng.integers(-5, 6, nR); S = rng.integers(-5, 6, nS)
        a, b = nested(R, S)
        p = rng.permutation(a.size)
        add(f"rand_small_{k}", R, S, a[p], b[p], "random small keys, shuffled result")
    R = np.array([-2**31, 2**31 - 1, -1, 0, -1, 2**31 - 1]); S = np.array([-1, 2**31 - 1, -2**31, 7, -1])
    a, b = nested(R, S)
    add("extreme_keys", R, S, a, b, "INT32_MIN / INT32_MAX / -1 (0xFFFFFFFF) are legal keys: the reference reserves none")
    R = rng.permutation(1024).astype(np.int32); S = rng.integers(0, 2048, 4096).astype(np.int32)
    a, b = nested(R, S)
    add("c1_shape_1k_x_4k", R, S, a, b, "BASELINE.json config 1 shape: unique build, ~50 % hits")
    swapped = (b.copy(), a.copy())
    add("c1_shape_swapped_columns", R, S, swapped[0], swapped[1], "columns swapped -> 0")

    out = Path(__file__).with_name("check_verdicts.json")
    out.write_text(json.dumps(dict(source="oracle/_ref/shared.so:check (shared_stuff/shared.cpp:129-172, g++ -O2, unmodified)", cases=cases)))
    print(f"wrote {out} ({len(cases)} cases): verdicts", [c["verdict"] for c in cases])


if __name__ == "__main__":
    main()
