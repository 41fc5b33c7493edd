"""world_size-2 (and 4) CPU test of the multi-rank path with torch.distributed `gloo`: every rank builds its Decomp2D
block with the library's host code, exchanges halos with the library's own halo plan (pack list / recv slots / peers,
the lists the NCCL path uses on the GPU) through real point-to-point messages, evaluates the residual and Jacobian of
its block with the emulated device functions, and the gathered result must equal the 1-rank oracle bit for bit.
A distributed dot product (local partial + all_reduce) is checked against the global one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from cases import PAR_INDEX as P

PARS = dict(cases.DEFAULT_PARS, NLES=1.0)
CASES = {"gateway16": cases.gateway16, "gateway16_balanced": lambda **kw: cases.gateway16(balance=1, **kw), "box_np": lambda **kw: cases.box(6, 7, 4, False, seed=2, land_frac=0.3, **kw),
         "box_p": lambda **kw: cases.box(9, 8, 3, True, seed=7, land_frac=0.25, **kw)}


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, name, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu.emu import EmuTHCM
    s, landm = CASES[name](rank=rank, nranks=world)
    e = EmuTHCM(s, landm)
    for k, v in PARS.items():
        e.setpar(P[k], v)
    sg, _ = CASES[name]()
    x = cases.random_state(sg, landm, scale=0.3, zero_on_land=False)
    gid = e.local_gids()
    xl = x[gid].copy()
    # ---- halo exchange with the library's plan ----
    peers, send_idx, recv_slot = e.plan()
    halo = np.full(max(e.nhalo, 1), np.nan)
    xc = xl.reshape(-1, 6)
    reqs, rbufs = [], []
    for (q, so, sc, ro, rc) in peers:
        if sc:
            sbuf = torch.from_numpy(np.ascontiguousarray(xc[send_idx[so:so + sc]]).reshape(-1))
            reqs.append(dist.isend(sbuf, dst=int(q)))
        if rc:
            rb = torch.empty(rc * 6, dtype=torch.float64)
            rbufs.append((rb, ro, rc))
            reqs.append(dist.irecv(rb, src=int(q)))
    for r in reqs:
        r.wait()
    hview = halo.reshape(-1, 6)
    for rb, ro, rc in rbufs:
        hview[recv_slot[ro:ro + rc]] = rb.numpy().reshape(-1, 6)
    B = e.rhs(xl, halo)
    val = e.jacobian(xl, halo)
    rp, col = e.graph()
    # distributed dot: local partial + all_reduce (what dot_dev + ncclAllReduce do)
    part = torch.tensor([float(xl @ xl)], dtype=torch.float64)
    dist.all_reduce(part)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), gid=gid, B=B, val=val, rp=rp, col=col, hg=e.halo_gids(), dot=part.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("gateway16", 2), ("box_np", 2), ("box_p", 2), ("gateway16", 4), ("gateway16_balanced", 4)])
def test_two_rank_gloo_matches_global_oracle(name, world, tmp_path):
    from oracle.oracle import OracleTHCM
    port = free_port()
    mp.spawn(worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    s, landm = CASES[name]()
    o = OracleTHCM(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
    x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
    Bo = o.rhs(x)
    vo, _ = o.jacobian_graph(x)
    ro, co = o.graph()
    seen = np.zeros(o.ndim, int)
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        gid = d["gid"]
        seen[gid] += 1
        assert np.array_equal(d["B"], Bo[gid])
        rp, col, val, hg = d["rp"], d["col"], d["val"], d["hg"]
        nloc = len(gid)
        gcol = np.where(col < nloc, gid[np.minimum(col, nloc - 1)], hg[np.maximum(col - nloc, 0)])
        want_idx = np.concatenate([np.arange(ro[g], ro[g + 1]) for g in gid])
        assert np.array_equal(gcol, co[want_idx]) and np.array_equal(val, vo[want_idx])
        assert abs(d["dot"][0] - float(x @ x)) <= 1e-12 * float(x @ x)
    assert np.all(seen == 1)
