"""CPU tests (no GPU) of the host side of the decomposed path: the block-decomposition rule, and the halo message
schedule libpanslbm_b200.so computes on the host (pl_halo_describe: peers, population sets, region geometry, issue order).

The schedule is exercised for real: world_size-2 and -4 `gloo` process groups exchange numpy halos following it — every rank
posts its sends/receives in the library's issue order, exactly as the NCCL path does — and the assembled result must equal
the global periodic Stream / iStream (the invariant of the reference's MPI build, d3q15.h:257-600).  The per-site receive
logic restated here in numpy is the one of the CUDA pull (csrc/lbm_halo.cuh: pull_halo)."""
import os
import socket

import numpy as np
import pytest

C3 = np.array([[0, 1, 0, 0, -1, 0, 0, 1, -1, 1, 1, -1, 1, -1, -1],
               [0, 0, 1, 0, 0, -1, 0, 1, 1, -1, 1, -1, -1, 1, -1],
               [0, 0, 0, 1, 0, 0, -1, 1, 1, 1, -1, -1, -1, -1, 1]])
C2 = np.array([[0, 1, 0, -1, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1], [0]*9])


def split(l, m, pe):
    """reference block decomposition (d3q15.h:29-35)"""
    n = (l + pe)//m
    off = pe*n if m - pe > l % m else l - (m - pe)*n
    return n, off


def test_block_decomposition_covers_the_domain():
    for l in (7, 8, 9, 31, 81, 161):
        for m in (1, 2, 3, 4, 5):
            parts = [split(l, m, pe) for pe in range(m)]
            assert sum(n for n, _ in parts) == l
            pos = 0
            for n, off in parts:
                assert off == pos and n > 0
                pos += n


def test_message_sets_match_the_reference_tables():
    """SURVEY appendix A / d3q15.h:299-312, 361-363, 428-435: 5 per face site, 2 per edge site, 1 per corner"""
    import panslbm2_b200 as pl
    msgs = {d["o"]: d for d in pl.halo_describe(3, 8, 8, 8, 0, 2, 2, 2, False)}
    assert len(msgs) == 26
    assert msgs[(-1, 0, 0)]["pops"] == [4, 8, 11, 13, 14] and msgs[(1, 0, 0)]["pops"] == [1, 7, 9, 10, 12]
    assert msgs[(0, -1, 0)]["pops"] == [5, 9, 11, 12, 14] and msgs[(0, 0, 1)]["pops"] == [3, 7, 8, 9, 14]
    assert msgs[(0, -1, -1)]["pops"] == [11, 12] and msgs[(-1, -1, -1)]["pops"] == [11] and msgs[(1, 1, 1)]["pops"] == [7]
    assert msgs[(-1, 0, 0)]["rsize"] == 16 and msgs[(0, -1, -1)]["rsize"] == 4 and msgs[(1, 1, 1)]["rsize"] == 1
    inv = {d["o"]: d for d in pl.halo_describe(3, 8, 8, 8, 0, 2, 2, 2, True)}
    assert inv[(-1, 0, 0)]["pops"] == [1, 7, 9, 10, 12]          # iStream swaps the sets (d3q15.h:663-667)
    # undecomposed axes are not exchanged; D2Q9: 3 per edge site, 1 per corner (d2q9.h:193-216)
    assert sorted(d["o"] for d in pl.halo_describe(3, 8, 8, 8, 1, 2, 1, 1)) == [(-1, 0, 0), (1, 0, 0)]
    m2 = {d["o"]: d for d in pl.halo_describe(2, 9, 8, 1, 3, 2, 2, 1)}
    assert len(m2) == 8 and m2[(1, 0, 0)]["pops"] == [1, 5, 8] and m2[(-1, -1, 0)]["pops"] == [7]


# ---------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slot(CC, c, mask):
    return sum(1 for d in range(c) if all(CC[a][d] == CC[a][c] for a in range(3) if mask[a]))


def _rank_stream(rank, world, port, dim, size, m, inverse, out_dir):
    import torch
    import torch.distributed as dist
    import panslbm2_b200 as pl
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    CC = C3 if dim == 3 else C2
    nc = CC.shape[1]
    lx, ly, lz = size
    pe = (rank % m[0], (rank//m[0]) % m[1], rank//(m[0]*m[1]))
    (nx, ox), (ny, oy), (nz, oz) = split(lx, m[0], pe[0]), split(ly, m[1], pe[1]), split(lz, m[2], pe[2])
    n = (nx, ny, nz)
    G = np.random.RandomState(5).uniform(size=(nc, lz, ly, lx))          # the same global field on every rank
    loc = G[:, oz:oz + nz, oy:oy + ny, ox:ox + nx].copy()
    s = -1 if inverse else 1
    for _ in range(2):
        flat = loc.reshape(nc, -1)
        msgs = pl.halo_describe(2 if dim == 2 else 3, lx, ly, lz, rank, *m, inverse)
        by_code = {d["code"]: d for d in msgs}
        # pack (k_halo_pack) and post in the library's issue order: send message `code`, receive the one of the opposite side
        reqs, recv = [], {}
        for d in msgs:
            t = np.arange(d["rsize"])
            sites = d["base"] + (t % d["n1"])*d["s1"] + (t//d["n1"])*d["s2"]
            buf = torch.from_numpy(np.ascontiguousarray(np.stack([flat[c][sites] for c in d["pops"]])))
            src = by_code[d["recv_code"]]
            rb = torch.empty(d["npop"], d["rsize"], dtype=torch.float64)
            recv[d["recv_code"]] = rb
            reqs.append(dist.isend(buf, d["peer"]))
            reqs.append(dist.irecv(rb, src["peer"]))
        for r in reqs:
            r.wait()
        # local periodic stream, then the halo-aware pull for sources beyond a decomposed face (pull_halo)
        new = np.stack([np.roll(loc[c], (s*CC[2][c], s*CC[1][c], s*CC[0][c]), axis=(0, 1, 2)) for c in range(nc)])
        kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        co = (ii, jj, kk)
        for c in range(1, nc):
            dvec = [-s*CC[a][c] for a in range(3)]                       # where the source lies
            cross = [np.zeros_like(ii, dtype=bool) for _ in range(3)]
            for a in range(3):
                if dvec[a] != 0 and m[a] > 1:
                    cross[a] = co[a] == (0 if dvec[a] < 0 else n[a] - 1)
            anyc = cross[0] | cross[1] | cross[2]
            for site in zip(*np.nonzero(anyc)):
                k, j, i = site
                mask = [bool(cross[a][site]) for a in range(3)]
                o = tuple(dvec[a] if mask[a] else 0 for a in range(3))
                code = (o[0] + 1) + 3*(o[1] + 1) + 9*(o[2] + 1)
                sc = [(x + dvec[a]) % n[a] for a, x in enumerate((i, j, k))]
                ridx, rs = 0, 1
                for a in range(3):
                    if not mask[a] and not (dim == 2 and a == 2):
                        ridx += sc[a]*rs
                        rs *= n[a]
                new[c][site] = recv[code][_slot(CC, c, mask), ridx]
        loc = new
    np.save(os.path.join(out_dir, f"r{rank}.npy"), loc)
    np.save(os.path.join(out_dir, f"o{rank}.npy"), np.array([ox, oy, oz, nx, ny, nz]))
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,size,m,inverse", [(3, (7, 6, 5), (2, 1, 1), False), (3, (7, 6, 5), (2, 1, 1), True), (3, (7, 6, 6), (2, 2, 1), False),
                                                (3, (5, 4, 7), (1, 2, 2), True), (2, (9, 7, 1), (2, 2, 1), False)])
def test_gloo_exchange_following_the_schedule_equals_global_stream(tmp_path, dim, size, m, inverse):
    import torch.multiprocessing as mp
    world = m[0]*m[1]*m[2]
    mp.spawn(_rank_stream, args=(world, _free_port(), dim, size, m, inverse, str(tmp_path)), nprocs=world, join=True)
    CC = C3 if dim == 3 else C2
    nc = CC.shape[1]
    lx, ly, lz = size
    want = np.random.RandomState(5).uniform(size=(nc, lz, ly, lx))
    s = -1 if inverse else 1
    for _ in range(2):
        want = np.stack([np.roll(want[c], (s*CC[2][c], s*CC[1][c], s*CC[0][c]), axis=(0, 1, 2)) for c in range(nc)])
    got = np.zeros_like(want)
    for r in range(world):
        ox, oy, oz, nx, ny, nz = np.load(tmp_path / f"o{r}.npy")
        got[:, oz:oz + nz, oy:oy + ny, ox:ox + nx] = np.load(tmp_path / f"r{r}.npy")
    assert np.array_equal(got, want)
