"""Multi-GPU sharding: BGZF virtual-offset shards cut at BAI linear-index record starts, one process
per GPU, boundary counters exchanged with NCCL send/recv inside the library.  Needs >= 2 GPUs."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import sambamba_b200 as sb
        return sb.load_library().bdepth_device_count()
    except Exception:
        return 0


def _exome():
    import random
    rnd = random.Random(5)
    big = os.environ.get("BDEPTH_EMULATE") != "1"
    la, lc = (2000000, 1500000) if big else (200000, 150000)
    na, nc = (25, 15) if big else (4, 2)       # the small file of the emulation has few blocks: keep most of them outside the query
    return sorted({(0, s, s + 200) for s in (rnd.randrange(0, la - 300) for _ in range(na))} | {(2, s, s + 350) for s in (rnd.randrange(0, lc - 400) for _ in range(nc))})


def _rank_main(rank, world, path, uid, mode, q, tuning=None):
    try:
        sys.path.insert(0, helpers.ROOT)
        import sambamba_b200 as sb
        with sb.BDepth(path, device=rank) as b:
            if mode.startswith("index+"):       # input without .bai: every rank builds the index itself, then the shards are cut from it
                assert not b.has_index
                b.build_index()
                mode = mode[6:]
            b.set_shard(rank, world, uid)
            if tuning:
                b.set_tuning(*tuning)
            if mode.endswith("-m"):
                b.set_fix_mates(True)
                mode = mode[:-2]
            if mode == "base":
                got = b.run_base()
                st = b.stats()
                lo, hi = st["own_lo"], st["own_hi"]
                q.put((rank, "ok", lo, hi, got[:, lo:hi].copy(), st))
            elif mode == "exome":
                rows = b.run_regions(_exome(), [2, 8])
                q.put((rank, "ok", 0, 0, rows, b.stats()))
            elif mode == "regions":
                rows = b.run_regions([(0, 100, 9000), (0, 9000, 9100), (0, 60000, 140000), (2, 5, 100000)], [2, 8])
                q.put((rank, "ok", 0, 0, rows, b.stats()))
            else:
                rows = b.run_windows(1000, 0, [1, 10])
                q.put((rank, "ok", 0, 0, rows, b.stats()))
    except Exception as e:  # pragma: no cover
        q.put((rank, "err", 0, 0, repr(e), None))


def _run(world, path, mode, tuning=None, expect_errors=False):
    import sambamba_b200 as sb
    uid = sb.nccl_unique_id()
    if os.environ.get("BDEPTH_EMULATE") == "1":
        # CPU emulation of the pipeline (tests/emul): ranks are threads of this process, NCCL is a rendezvous between them
        import queue
        import threading
        q = queue.Queue()
        ts = [threading.Thread(target=_rank_main, args=(r, world, path, uid, mode, q, tuning)) for r in range(world)]
        for t in ts:
            t.start()
        res = [q.get(timeout=1500) for _ in range(world)]
        for t in ts:
            t.join(timeout=60)
        res.sort(key=lambda r: r[0])
        for r in res:
            assert expect_errors or r[1] == "ok", r
        return res
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank_main, args=(r, world, path, uid, mode, q, tuning)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    res.sort(key=lambda r: r[0])
    for r in res:
        assert expect_errors or r[1] == "ok", r
    return res


@pytest.fixture(scope="module")
def bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("multi")
    if os.environ.get("BDEPTH_EMULATE") == "1" and os.environ.get("BDEPTH_EMU_MULTI_FULL") != "1":       # same shape, a tenth of the size (the CPU emulation is slow; BDEPTH_EMU_MULTI_FULL=1: the hardware's size)
        return helpers.gen_bam(str(d / "m.bam"), "-r", "chrA:200000", "-r", "chrB:700", "-r", "chrC:150000", "-n", 30000, "-s", 11, "-t", 8)
    return helpers.gen_bam(str(d / "m.bam"), "-r", "chrA:2000000", "-r", "chrB:700", "-r", "chrC:1500000", "-n", 300000, "-s", 11, "-t", 8)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_base_equals_oracle(bam, world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    want, ost = helpers.oracle_counts(bam)
    res = _run(world, bam, "base")
    got = np.zeros_like(want)
    prev_hi = 0
    total_records = 0
    for rank, _, lo, hi, arr, st in res:
        assert lo == prev_hi or hi == lo, "owned ranges must tile the genome in rank order"
        got[:, lo:hi] = arr
        prev_hi = max(prev_hi, hi)
        total_records += st["n_records"]
    assert prev_hi == want.shape[1]
    assert total_records == ost.n_records, "every record belongs to exactly one shard"
    assert np.array_equal(got, want)
    assert sum(r[5]["covered_positions"] for r in res) == int((want.sum(axis=0) > 0).sum())
    # (plain shards exchange no counters since round 2: a rank reads the previous ranks' reads that reach into its positions itself -- the
    # zone the BAI linear index points to -- and delivers its positions while its shard is still streaming; the bit-exact comparison
    # above is what shows that the reads straddling a shard boundary were counted exactly once)
    assert all(r[5]["halo_bytes_sent"] == 0 for r in res)


@pytest.mark.parametrize("world,tuning", [(2, (1 << 20, 3)), (4, (1 << 20, 1)), (3, (0, 2))])
def test_sharded_with_small_batches_and_sub_batches(bam, world, tuning):
    """Several batches per rank and several sub-batches per batch (the shard limit then falls inside or before a sub-batch)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    want, ost = helpers.oracle_counts(bam)
    res = _run(world, bam, "base", tuning)
    got = np.zeros_like(want)
    for rank, _, lo, hi, arr, st in res:
        got[:, lo:hi] = arr
    assert sum(r[5]["n_records"] for r in res) == ost.n_records
    assert any(r[5]["n_batches"] > 1 for r in res) or tuning[0] == 0
    assert np.array_equal(got, want)


def test_sharded_windows_equal_single_gpu(bam):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import sambamba_b200 as sb
    with sb.BDepth(bam) as b:
        want = b.run_windows(1000, 0, [1, 10])
    res = _run(2, bam, "windows")
    for r in res:
        assert r[4] == want      # the statistics are all-reduced, so every rank holds the full table


@pytest.fixture(scope="module")
def pairs_bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("multim")
    small = os.environ.get("BDEPTH_EMULATE") == "1"
    return helpers.gen_bam(str(d / "mp.bam"), "-r", "chrA:%d" % (200000 if small else 2000000), "-r", "chrB:700", "-r", "chrC:%d" % (150000 if small else 1500000),
                           "-n", 30000 if small else 300000, "--pairs", 7, "-s", 12, "-t", 8)


@pytest.mark.parametrize("world,tuning", [(2, None), (4, None), (8, None), (3, (1 << 20, 0)), (2, (1 << 16, 0))])
def test_fix_mates_on_several_ranks(pairs_bam, world, tuning):
    """-m with shards: every rank reads a zone of its neighbours' records around its shard, a pair cut by a shard boundary is
    fixed by the rank that owns its first mate and the correction travels with the halo exchange."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    want, npc = helpers.oracle_counts_fix_mates(pairs_bam)
    res = _run(world, pairs_bam, "base-m", tuning)
    got = np.zeros_like(want)
    for rank, _, lo, hi, arr, st in res:
        got[:, lo:hi] = arr
    assert np.array_equal(got, want)
    assert sum(r[5]["mate_pair_columns"] for r in res) == npc and npc > 10000
    _, ost = helpers.oracle_counts(pairs_bam)
    assert sum(r[5]["n_records"] for r in res) == ost.n_records


def test_fix_mates_windows_and_regions_on_several_ranks(pairs_bam):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import sambamba_b200 as sb
    with sb.BDepth(pairs_bam) as b:
        b.set_fix_mates(True)
        want_w = b.run_windows(1000, 0, [1, 10])
        want_r = b.run_regions([(0, 100, 9000), (0, 9000, 9100), (0, 60000, 140000), (2, 5, 100000)], [2, 8])
    for world in [w for w in (2, 3) if _n_gpus() >= w]:
        for r in _run(world, pairs_bam, "windows-m"):
            assert r[4] == want_w
        for r in _run(world, pairs_bam, "regions-m"):
            assert r[4] == want_r


@pytest.mark.parametrize("world", [2, 3])
def test_scattered_regions_divide_their_chunks_among_ranks(bam, world):
    """BASELINE configs[4] (`depth region -L exome.bed` on several GPUs): only the BAI chunks of the regions are staged, and
    consecutive runs of them go to consecutive ranks; the region table is all-reduced."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import sambamba_b200 as sb
    with sb.BDepth(bam) as b:
        b.run_base(collect=False)
        total_blocks = b.stats()["n_blocks"]
        want = b.run_regions(_exome(), [2, 8])
        one = b.stats()["n_blocks"]
    res = _run(world, bam, "exome")
    for r in res:
        assert r[4] == want
    blocks = [r[5]["n_blocks"] for r in res]
    assert sum(blocks) <= one + world and sum(blocks) < total_blocks, (blocks, one, total_blocks)
    assert sum(1 for n in blocks if n) >= 2, blocks


def test_zone_blocks_whose_reads_all_end_before_the_shard(tmp_path):
    """The zone a rank re-reads begins at the first read that overlaps the 16 kbp window of its first own read (BAI linear index).  The
    reads behind that one need not reach the window: here a read with a long skip overlaps it, and the next BGZF members hold only short
    reads that end before it.  With one member per sub-batch such a sub-batch has nothing to count for this rank (hardware run of round 2:
    this case ended in "read extends past the end of the reference space")."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import random
    rnd = random.Random(7)
    M, N = 0, 3
    def seq():
        return "".join(rnd.choice("ACGT") for _ in range(40))
    reads = [(0, p, 60, 0, [(40, M)], seq(), "a%d" % i) for i, p in enumerate(sorted(rnd.randrange(0, 13000) for _ in range(300)))]
    reads.append((0, 14000, 60, 0, [(20, M), (2960, N), (20, M)], seq(), "skip"))
    reads += [(0, p, 60, 0, [(40, M)], seq(), "b%d" % i) for i, p in enumerate(sorted(rnd.randrange(14000, 15000) for _ in range(200)))]
    reads += [(0, p, 60, 0, [(40, M)], seq(), "c%d" % i) for i, p in enumerate(sorted(rnd.randrange(16400, 32000) for _ in range(100)))]
    reads += [(0, p, 60, 0, [(40, M)], seq(), "d%d" % i) for i, p in enumerate(sorted(rnd.randrange(32700, 60000) for _ in range(500)))]
    p = helpers.write_bam(str(tmp_path / "z.bam"), [("r0", 70000)], reads, block=4096, bins="auto", index=False)
    open(p + ".bai", "wb").write(helpers.oracle_build_bai(p))
    want, ost = helpers.oracle_counts(p)
    res = _run(2, p, "base", (1 << 16, 1))
    got = np.zeros_like(want)
    for rank, _, lo, hi, arr, st in res:
        got[:, lo:hi] = arr
    assert res[1][2] > 16384 and res[1][5]["n_records"] > 0, "the second rank owns the positions from its first read on, inside the window the zone was read for"
    assert sum(r[5]["n_records"] for r in res) == ost.n_records
    assert np.array_equal(got, want)


def test_unindexed_input_on_several_ranks(bam, tmp_path):
    """A BAM without .bai: each rank builds the index on its GPU (bdepth_build_index), adopts it and takes its shard of the file."""
    import shutil
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    p = str(tmp_path / "noidx.bam")
    shutil.copy(bam, p)
    want, ost = helpers.oracle_counts(bam)
    res = _run(2, p, "index+base")
    got = np.zeros_like(want)
    for rank, _, lo, hi, arr, st in res:
        got[:, lo:hi] = arr
    assert sum(r[5]["n_records"] for r in res) == ost.n_records and res[1][2] > 0
    assert np.array_equal(got, want)


def test_rank_without_a_passing_read_owns_nothing(tmp_path):
    """The middle shard holds only reads the default filter drops (MAPQ 0): that rank owns no position -- own_lo = own_hi = 0 -- while its
    counter window begins at its shard.  (Found by fuzzing under the emulation: the owned part of the window wrapped around and the
    covered-positions kernel ran over it.)"""
    if _n_gpus() < 3:
        pytest.skip("needs 3 GPUs")
    import random
    rnd = random.Random(3)
    def seq():
        return "".join(rnd.choice("ACGT") for _ in range(40))
    reads = [(0, p, 60, 0, [(40, 0)], seq(), "a%d" % i) for i, p in enumerate(sorted(rnd.randrange(0, 10000) for _ in range(150)))]
    reads += [(0, p, 0, 0, [(40, 0)], seq(), "b%d" % i) for i, p in enumerate(sorted(rnd.randrange(20000, 150000) for _ in range(900)))]      # both cuts fall in here, several 16 kbp windows wide
    reads += [(0, p, 60, 0, [(40, 0)], seq(), "c%d" % i) for i, p in enumerate(sorted(rnd.randrange(160000, 170000) for _ in range(150)))]
    p = helpers.write_bam(str(tmp_path / "f.bam"), [("r0", 180000)], reads, block=4096, bins="auto", index=False)
    open(p + ".bai", "wb").write(helpers.oracle_build_bai(p))
    want, ost = helpers.oracle_counts(p)
    res = _run(3, p, "base")
    got = np.zeros_like(want)
    for rank, _, lo, hi, arr, st in res:
        got[:, lo:hi] = arr
    assert (res[1][2], res[1][3]) == (0, 0) and res[1][5]["n_records"] > 0 and res[1][5]["n_records_pass"] == 0
    assert sum(r[5]["n_records"] for r in res) == ost.n_records
    assert np.array_equal(got, want)


def test_a_rank_that_refuses_takes_the_others_down_with_it(tmp_path):
    """-m on several ranks refuses a chain of same-name reads that reaches more than 64 BGZF members past a shard boundary -- on the rank that
    owns its leader only.  That rank still joins the boundary all-gather with a "failed" mark, so the other ranks end with an error of
    their own instead of waiting in the collective (found by the differential fuzzer: three such cases hung the run)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    n = 6000                     # 60 kbp of one name, every read overlapping the next four: the chain crosses the shard boundary (a 16 kbp window of the index) by far
    reads = [(0, 100 + 10 * i, 30, 0x1 | (0x40 if i % 2 == 0 else 0x80), [(50, 0)], "ACGTA" * 10, "chain") for i in range(n)]
    p = helpers.write_bam(str(tmp_path / "chain.bam"), [("r0", 70000)], reads, block=400, bins="auto", index=False)
    import sambamba_b200 as sb
    with sb.BDepth(p) as b:
        open(p + ".bai", "wb").write(b.build_index())
    res = _run(2, p, "base-m", expect_errors=True)
    assert [r[1] for r in res] == ["err", "err"], res
    msgs = sorted(r[4] for r in res)
    assert any("64 BGZF blocks past a shard boundary" in m for m in msgs) and any("stopped with an error" in m for m in msgs), msgs
    # ... and the same file without -m is simply counted
    want, _ = helpers.oracle_counts(p)
    res = _run(2, p, "base")
    got = np.zeros_like(want)
    for rank, status, lo, hi, arr, st in res:
        got[:, lo:hi] = arr
    assert np.array_equal(got, want)
