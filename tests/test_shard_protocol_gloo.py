"""World-size-2 check of the multi-GPU protocol on CPU (gloo): the shard plan from the library (host-only
bdepth_plan_shards) + a Python mirror of the boundary exchange (same ownership rule and send/recv pattern as
exchange_boundaries() in sambamba_b200/csrc/bdepth.cu) must reproduce the whole-file oracle counters."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _shard_counts(path, u_lo, u_hi):
    """Per-position counters of the records whose start offset lies in [u_lo, u_hi) (oracle walk in numpy/py)."""
    import struct
    u = helpers.oracle_inflate(path)
    first, refs = helpers.header_first_record_offset(u)
    lin0, lin = [], 0
    for _, L in refs:
        lin0.append(lin)
        lin += L
    counts = np.zeros((7, lin), np.uint32)
    b = u.tobytes()
    mn, mx, nrec = 2 ** 63, 0, 0
    nt5 = [4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4]
    for (o, ref, pos, flag, mapq, ncig, span) in helpers.parse_records(u, first):
        if not (u_lo <= o < u_hi):
            continue
        nrec += 1
        if ref < 0 or mapq == 0 or (flag & 0x604) or span == 0:
            continue
        lname = b[o + 12]
        lseq = struct.unpack_from("<i", b, o + 20)[0]
        cg = o + 36 + lname
        sq = cg + 4 * ncig
        g, q = lin0[ref] + pos, 0
        mn, mx = min(mn, g), max(mx, g + span)
        for k in range(ncig):
            c = struct.unpack_from("<I", b, cg + 4 * k)[0]
            l, op = c >> 4, c & 15
            if op in (0, 7, 8):
                for j in range(l):
                    byte = b[sq + ((q + j) >> 1)]
                    nib = (byte & 15) if ((q + j) & 1) else (byte >> 4)
                    counts[nt5[nib], g + j] += 1
                g += l
                q += l
            elif op in (2, 3):
                counts[5 if op == 2 else 6, g:g + l] += 1
                g += l
            elif op in (1, 4):
                q += l
    return counts, mn, mx, nrec


def _worker(rank, world, path, cuts_u, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = cuts_u[rank], cuts_u[rank + 1]
    counts, mn, mx, nrec = _shard_counts(path, lo, hi)
    # all-gather (min_start, max_end)
    mine = torch.tensor([mn, mx], dtype=torch.int64)
    allp = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allp, mine)
    mins = [int(t[0]) for t in allp]
    maxs = [int(t[1]) for t in allp]
    total = counts.shape[1]
    nonempty = [mins[r] != 2 ** 63 for r in range(world)]

    def own_hi(r):
        for s in range(r + 1, world):
            if nonempty[s]:
                return mins[s]
        return total

    def own_lo(r):
        return mins[r] if any(nonempty[:r]) else 0
    reqs, recvs = [], []
    if nonempty[rank]:
        for j in range(rank + 1, world):
            if nonempty[j] and mins[j] < mx:
                a, b = max(mins[j], mn), min(mx, own_hi(j))
                if a < b:
                    reqs.append(dist.isend(torch.from_numpy(counts[:, a:b].astype(np.int64).copy()), j))
        for i in range(rank):
            if nonempty[i] and maxs[i] > mins[rank]:
                a, b = max(mins[rank], mins[i]), min(maxs[i], own_hi(rank))
                if a < b:
                    buf = torch.zeros((7, b - a), dtype=torch.int64)
                    recvs.append((a, b, buf, dist.irecv(buf, i)))
    for a, b, buf, r in recvs:
        r.wait()
        counts[:, a:b] += buf.numpy().astype(np.uint32)
    for r in reqs:
        r.wait()
    a, b = (own_lo(rank), own_hi(rank)) if nonempty[rank] else (0, 0)
    q.put((rank, a, b, counts[:, a:b].copy(), nrec))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_reproduces_whole_file(tmp_path):
    import sambamba_b200 as sb
    p = helpers.gen_bam(str(tmp_path / "g.bam"), "-r", "chrA:60000", "-r", "chrB:40000", "-n", 6000, "-s", 9, "-t", 2)
    world = 2
    cuts = sb.plan_shards(p, world)
    # virtual offset -> inflated offset
    raw = open(p, "rb").read()
    off, uoff, table = 0, 0, {}
    while off + 18 <= len(raw):
        bs = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        isz = int.from_bytes(raw[off + bs - 4:off + bs], "little")
        if isz == 0:
            break
        table[off] = uoff
        uoff += isz
        off += bs
    cuts_u = [0] + [table[c >> 16] + (c & 0xFFFF) for c in cuts] + [uoff]
    assert 0 < cuts_u[1] < uoff, "the test file must actually be split"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, p, cuts_u, port, q)) for r in range(world)]
    for x in ps:
        x.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for x in ps:
        x.join(timeout=30)
    want, ost = helpers.oracle_counts(p)
    got = np.zeros_like(want)
    prev = 0
    for rank, a, b, arr, nrec in res:
        assert a == prev
        got[:, a:b] = arr
        prev = b
    assert prev == want.shape[1]
    assert sum(r[4] for r in res) == ost.n_records
    assert np.array_equal(got, want)
