"""Host-side helpers of the C library that need no GPU: grid digests, the boxed-output expander, the multi-GPU
partitions, and the millimetre form of the synthetic ranges."""
import ctypes as C

import numpy as np
import pytest


def _numpy_digest(g):
    m = np.uint64
    with np.errstate(over="ignore"):
        z = ((np.arange(g.size, dtype=m) << m(8)) | g.reshape(-1).view(np.uint8).astype(m)) + m(0x9E3779B97F4A7C15)
        z = (z ^ (z >> m(30))) * m(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> m(27))) * m(0x94D049BB133111EB)
        return int((z ^ (z >> m(31))).sum(dtype=m))


def test_grid_digest_host_equals_numpy_and_the_oracles_restatement(pkg, oracle):
    rng = np.random.default_rng(2)
    for shape in ((1, 1), (7, 13), (400, 400)):
        g = rng.integers(-128, 128, shape, dtype=np.int8)
        assert pkg.grid_hash(g) == _numpy_digest(g) == oracle.grid_hash64(g)
    z = np.zeros((50, 50), np.int8)
    assert pkg.grid_hash(z) != 0                      # zero cells still contribute: the digest is position-sensitive
    z2 = z.copy(); z2[3, 4] = 1
    z3 = z.copy(); z3[4, 3] = 1
    assert len({pkg.grid_hash(z), pkg.grid_hash(z2), pkg.grid_hash(z3)}) == 3


def test_unpack_boxed_expands_boxes_and_rejects_bad_ones(pkg):
    p = pkg.make_params(40, 30, 0.1)
    rng = np.random.default_rng(3)
    boxes = np.array([[4, 2, 16, 9], [0, 0, 0, 0], [36, 20, 40, 30]], np.int32)
    packed = rng.integers(-80, 81, 1000, dtype=np.int8)
    offsets = np.array([16, 500, 300], np.uint64)
    dense = pkg.unpack_boxed(p, boxes, offsets, packed)
    want = np.zeros((3, 30, 40), np.int8)
    want[0, 2:9, 4:16] = packed[16:16 + 7 * 12].reshape(7, 12)
    want[2, 20:30, 36:40] = packed[300:300 + 10 * 4].reshape(10, 4)
    assert np.array_equal(dense, want) and not dense[1].any()
    bad = boxes.copy(); bad[0, 2] = 41
    with pytest.raises(pkg.UqsError) as e:
        pkg.unpack_boxed(p, bad, offsets, packed)
    assert e.value.code == pkg.ERR_BAD_ARG


def test_partitions_cover_everything_exactly_once(pkg):
    for n, world in ((4096, 8), (13, 4), (3, 8), (1024, 3)):
        seen = []
        for r in range(world):
            first, cnt = pkg.flight_shard(n, r, world)
            seen += list(range(first, first + cnt))
        assert seen == list(range(n))
        sizes = [pkg.flight_shard(n, r, world)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    for H, world in ((16384, 8), (400, 3), (2000, 7), (10, 4)):
        rows = []
        for r in range(world):
            r0, cnt = pkg.row_band(H, r, world)
            assert r0 % 4 == 0 or r0 == H
            rows += list(range(r0, r0 + cnt))
        assert rows == list(range(H))
    with pytest.raises(ValueError):
        pkg.flight_shard(10, 4, 4)


def test_millimetre_form_of_the_synthetic_ranges_is_lossless(synth):
    w = synth.on_mm_lattice(synth.scaled(synth.CONFIGS["c3"], n_flights=3, n_samples=500))
    d = synth.generate(w)
    mm = synth.ranges_to_mm(d["ranges"])
    assert mm.dtype == np.uint16 and (mm == 0xFFFF).sum() == np.isnan(d["ranges"]).sum() > 0
    assert np.array_equal(synth.mm_to_ranges(mm).view(np.uint32), d["ranges"].view(np.uint32))
    # the float log without the lattice is NOT representable: the flag matters
    d0 = synth.generate(synth.scaled(synth.CONFIGS["c3"], n_flights=3, n_samples=500))
    assert not np.array_equal(synth.mm_to_ranges(synth.ranges_to_mm(d0["ranges"])).view(np.uint32), d0["ranges"].view(np.uint32))
    # everything but the ranges is the same log
    assert np.array_equal(d0["of_rate_x"], d["of_rate_x"]) and np.array_equal(d0["yaw_deg"], d["yaw_deg"])


def test_bench_digest_of_a_sharded_job_is_the_digest_of_the_whole_job():
    """bench.py reduces per-rank digests with a sum mod 2^64 (hash) and compares with the same job on one GPU
    (hash_n1): the combination must not depend on how the grids are split over ranks."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(9)
    h = rng.integers(0, 2 ** 63, 1000, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    whole = bench.combine_hashes(h, 0)
    for cuts in ((0, 500, 1000), (0, 1, 999, 1000), (0, 125, 250, 375, 500, 625, 750, 875, 1000)):
        parts = sum(bench.combine_hashes(h[a:b], a) for a, b in zip(cuts, cuts[1:])) & bench.M64
        assert parts == whole
    assert bench.combine_hashes(h[::-1].copy(), 0) != whole          # which grid belongs to which flight matters
    # the four 16-bit limbs an int64 all-reduce carries reassemble the sum exactly
    vals = [int(v) for v in h[:8]]
    limbs = [sum((v >> (16 * i)) & 0xFFFF for v in vals) for i in range(4)]
    assert sum(l << (16 * i) for i, l in enumerate(limbs)) & bench.M64 == sum(vals) & bench.M64
