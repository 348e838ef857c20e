"""N > 1 host logic on CPU (gloo, world_size 2): contiguous sharding of the cell index space and the
single all-reduce of the Mu sums.  The per-cell results come from the CPU oracle here (no GPU in this
container); on the GPU box bench.py runs the same logic over NCCL with the CUDA engine."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ncells_per_rank, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from fatode_b200.inputs import perturbed_states
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f01 = float(np.float32(0.1))
    rtol, atol = np.full(32, f01 * 1e-4), np.full(32, f01 * 1e-6)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5
    t0 = 43200.0
    # rank r owns global cells [r*n, (r+1)*n): the generator is counter based, so no data is exchanged
    y0 = perturbed_states(O.initial_state("cbm4"), ncells_per_rank, first_cell=rank * ncells_per_rank)
    r = O.ros_adj("cbm4", y0, 32, t0, t0 + 2 * 3600.0, atol, rtol, ic, nthreads=2)
    musum = torch.from_numpy(r["mu"].sum(axis=0).copy())
    dist.all_reduce(musum)                      # the path's only collective
    nacc = torch.tensor([int(r["istatus"][:, 3].sum())])
    dist.all_reduce(nacc)
    if rank == 0:
        out.put((musum.numpy(), int(nacc.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mu_sum_equals_single_process_sum():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from fatode_b200.inputs import perturbed_states
    world, n = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    musum, nacc = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f01 = float(np.float32(0.1))
    rtol, atol = np.full(32, f01 * 1e-4), np.full(32, f01 * 1e-6)
    ic = np.zeros(20, dtype=np.int32)
    ic[2] = 5
    t0 = 43200.0
    y0 = perturbed_states(O.initial_state("cbm4"), world * n)
    ref = O.ros_adj("cbm4", y0, 32, t0, t0 + 2 * 3600.0, atol, rtol, ic)
    s = ref["mu"].sum(axis=0)
    absum = np.abs(ref["mu"]).sum(axis=0)
    assert nacc == int(ref["istatus"][:, 3].sum())
    assert np.all(np.abs(musum - s) <= 1e-12 * absum + 1e-300)


def _erk_worker(rank, world, port, per_rank, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from fatode_b200.inputs import swe_members
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = float(np.float32(2.1573758800627289e-003))
    tol = np.full(4800, 1.0e-3)
    # rank r owns ensemble members [r*n, (r+1)*n): counter-based generator, no data path collective for this family
    y0 = swe_members(per_rank, first_member=rank * per_rank)
    r = O.erk_adj("swe", y0, 1, 0.0, 3 * dt, tol, tol, nthreads=2)
    mine = torch.from_numpy(np.concatenate([r["lambda_"][:, 0, :8].ravel(), r["y"][:, :8].ravel(), r["istatus"][:, :5].ravel().astype(np.float64)]))
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)                 # only to compare in the test: the path itself exchanges nothing
    if rank == 0:
        out.put(np.stack([p.numpy() for p in parts]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_shallow_water_adjoint_equals_single_process():
    """ERK_ADJ over an ensemble split across two ranks (no collective on the data path): the union of the shards equals one run."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from fatode_b200.inputs import swe_members
    world, n = 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_erk_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dt = float(np.float32(2.1573758800627289e-003))
    tol = np.full(4800, 1.0e-3)
    ref = O.erk_adj("swe", swe_members(world * n), 1, 0.0, 3 * dt, tol, tol)
    for r in range(world):
        sl = slice(r * n, (r + 1) * n)
        want = np.concatenate([ref["lambda_"][sl, 0, :8].ravel(), ref["y"][sl, :8].ravel(), ref["istatus"][sl, :5].ravel().astype(np.float64)])
        assert np.array_equal(parts[r], want)
