"""Multi-GPU parity (needs >= 2 CUDA devices; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
one process per GPU, NCCL halo exchange + dot all-reduces.  The decomposed CUDA path must reproduce the 1-rank oracle:
residual and Jacobian bit-exact on every block, SpMV <= 1e-13, GMRES history equal to the 1-GPU history to 1e-10
(the multi-GPU path reproduces the GLOBAL answer on any GPU count -- cf. test_matrix.C:156-200 which asserts 1e-8)."""
import os
import socket

import numpy as np
import pytest

import cases
from cases import PAR_INDEX as P

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

PARS = dict(cases.DEFAULT_PARS, NLES=1.0)
CASES = {"gateway16": cases.gateway16, "global4deg": cases.global4deg,
         "global4deg_balanced": lambda **kw: cases.global4deg(balance=1, **kw),     # ocean-weighted cut lines
         "gateway16_balanced": lambda **kw: cases.gateway16(balance=1, **kw),
         "box_np": lambda **kw: cases.box(12, 10, 4, False, seed=2, land_frac=0.3, **kw)}


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, name, outdir):
    import torch.distributed as dist
    import iemic_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s, landm = CASES[name](rank=rank, nranks=world, device=rank)
    t = iemic_b200.THCM(s, landm, dist.group.WORLD)
    for k, v in PARS.items():
        t.setParameter(k, v)
    sg, _ = CASES[name]()
    x = cases.consistent_state(sg, landm, scale=0.1)
    gid = t.local_gids()
    xd = torch.from_numpy(x[gid]).cuda()
    F = t.new_vector()
    t.evaluate(xd, F, True)
    val = t.jacobian_values_host()
    rp, col = t.graph()
    v = np.random.default_rng(5).standard_normal(len(x))
    y = t.new_vector()
    t.applyMatrix(torch.from_numpy(v[gid]).cuda(), y)
    nrm = t.norm(F)
    t.buildPreconditioner(1)
    sol = t.new_vector()
    res, hist = t.gmres(F, sol, tol=1e-8, maxit=30, restart=15)
    # batched DGKS orthogonalisation (multi_dot with the fused LL all-reduce over peer memory)
    sol2 = t.new_vector()
    res2, hist2 = t.gmres(F, sol2, tol=1e-8, maxit=30, restart=15, ortho="dgks")
    # the same solve on full-length vectors (the default above ran on the ocean cells only, halo through the LL slots)
    sol3 = t.new_vector()
    res3, hist3 = t.gmres(F, sol3, tol=1e-8, maxit=30, restart=15, ortho="dgks", full_space=True)
    # back-to-back operator applications with no reduction in between: the two alternating halo buffers of the P2P push
    y2, y3 = t.new_vector(), t.new_vector()
    for _ in range(5):
        t.applyMatrix(y, y2)
        t.applyMatrix(y2, y3)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), gid=gid, F=F.cpu().numpy(), val=val, rp=rp, col=col, hg=t.halo_gids(),
             y=y.cpu().numpy(), nrm=nrm, hist=hist, sol=sol.cpu().numpy(), iters=res.iters, hist2=hist2, iters2=res2.iters,
             y3=y3.cpu().numpy(), hist3=hist3, iters3=res3.iters, sol2=sol2.cpu().numpy(), sol3=sol3.cpu().numpy())
    t.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", list(CASES))
def test_multi_gpu_reproduces_global_answer(name, tmp_path):
    import torch.multiprocessing as mp
    from oracle.oracle import OracleTHCM, spmv
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = int(os.environ.get("THCM_TEST_WORLD", "0")) or (8 if ngpu >= 8 else 4 if ngpu >= 4 else 2)
    mp.spawn(worker, args=(world, free_port(), name, str(tmp_path)), nprocs=world, join=True)
    s, landm = CASES[name]()
    o = OracleTHCM(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
    x = cases.consistent_state(s, landm, scale=0.1)
    Fo = -o.rhs(x)
    vo, _ = o.jacobian_graph(x)
    ro, co = o.graph()
    v = np.random.default_rng(5).standard_normal(len(x))
    yo = spmv(ro, co, vo, v)
    # single-GPU reference history
    import iemic_b200
    t1 = iemic_b200.THCM(s, landm)
    for k, vv in PARS.items():
        t1.setParameter(k, vv)
    F1 = t1.new_vector()
    t1.evaluate(torch.from_numpy(x).cuda(), F1, True)
    t1.buildPreconditioner(1)
    sol1 = t1.new_vector()
    res1, hist1 = t1.gmres(F1, sol1, tol=1e-8, maxit=30, restart=15)
    res1d, hist1d = t1.gmres(F1, t1.new_vector(), tol=1e-8, maxit=30, restart=15, ortho="dgks")
    y3o = spmv(ro, co, vo, spmv(ro, co, vo, yo))
    seen = np.zeros(o.ndim, int)
    y_all = np.zeros(o.ndim)
    y3_all = np.zeros(o.ndim)
    sol2_all = np.zeros(o.ndim); sol3_all = np.zeros(o.ndim)
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        gid = d["gid"]
        seen[gid] += 1
        assert np.array_equal(d["F"], Fo[gid])
        want_idx = np.concatenate([np.arange(ro[g], ro[g + 1]) for g in gid])
        assert np.array_equal(d["val"], vo[want_idx])
        y_all[gid] = d["y"]
        assert abs(d["nrm"] - np.linalg.norm(Fo)) <= 1e-13 * np.linalg.norm(Fo)
        k = min(len(hist1), len(d["hist"]))
        assert abs(int(d["iters"]) - res1.iters) <= 1
        assert np.abs(d["hist"][:k] - hist1[:k]).max() <= 1e-10
        k = min(len(hist1d), len(d["hist2"]))
        assert abs(int(d["iters2"]) - res1d.iters) <= 1
        assert np.abs(d["hist2"][:k] - hist1d[:k]).max() <= 1e-10
        y3_all[gid] = d["y3"]
        # ocean-only Krylov space (default) against full-length vectors on the same ranks
        k = min(len(d["hist2"]), len(d["hist3"]))
        assert abs(int(d["iters2"]) - int(d["iters3"])) <= 1 and np.abs(d["hist2"][:k] - d["hist3"][:k]).max() <= 1e-10
        sol2_all[gid] = d["sol2"]; sol3_all[gid] = d["sol3"]
    assert np.all(seen == 1)
    assert np.linalg.norm(sol2_all - sol3_all) <= 1e-9 * np.linalg.norm(sol3_all) and np.linalg.norm(sol3_all) > 0
    assert np.linalg.norm(y3_all - y3o) <= 1e-12 * np.linalg.norm(y3o)
    assert np.linalg.norm(y_all - yo) <= 1e-13 * np.linalg.norm(yo)
    t1.close()
