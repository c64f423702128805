"""CPU, gloo, world_size 2: the data-parallel step (shard -> per-rank backward -> ONE all-reduce -> Adam) leaves every
replica with identical parameters, equal to a single process applying the mean gradient of the same views.
Per-rank compute is the CPU oracle (no GPU here); the reduce / shard / buffer logic is the code the NCCL path uses."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LRS = [0.00032, 0.0025, 0.0025 / 20, 0.05, 0.005, 0.001]


def _rank_gradients(rank, P):
    """Raw-parameter gradients of view `rank` of a tiny shared scene, from the C oracle (+ its activation backward)."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import ctypes as C
    import oracle_c
    import photo_slam_b200.synthetic as syn
    cam0 = syn.make_camera(96, 64, 80.0, 80.0)
    sc = syn.make_scene(P, cam0, seed=5, scale_px=5.0)
    R, t = syn.random_pose(np.random.default_rng(10 + rank), 0.1, 0.1)
    cam = syn.make_camera(96, 64, 80.0, 80.0, R, t)
    act = syn.activate(sc)
    f = oracle_c.forward(cam, act)
    gt = syn.target_image(64, 96, seed=rank)
    _, _, _, dpix = oracle_c.loss(f["out_color"], gt)
    b = oracle_c.backward(cam, act, f, dpix)
    L = oracle_c.lib()
    g_op, g_sc, g_rot = np.zeros((P, 1), np.float32), np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32)
    L.orc_activations_backward(C.c_int(P), oracle_c._p(sc["opacity"]), oracle_c._p(sc["scaling"]), oracle_c._p(sc["rotation"]),
                               oracle_c._p(b["dL_dopacity"]), oracle_c._p(b["dL_dscale"]), oracle_c._p(b["dL_drot"]), oracle_c._p(g_op),
                               oracle_c._p(g_sc), oracle_c._p(g_rot))
    grads = [b["dL_dmean3D"], b["dL_dsh"][:, :1, :], b["dL_dsh"][:, 1:, :], g_op, g_sc, g_rot]
    params = [sc["xyz"], sc["features_dc"], sc["features_rest"], sc["opacity"], sc["scaling"], sc["rotation"]]
    return params, grads, oracle_c


def _adam_all(params, grads, oracle_c, scale):
    out = []
    for p, g, lr in zip(params, grads, LRS):
        pn, _, _ = oracle_c.adam(p.ravel(), (np.asarray(g, np.float32) * np.float32(scale)).ravel(), np.zeros(p.size, np.float32),
                                 np.zeros(p.size, np.float32), lr, 1)
        out.append(pn)
    return out


def _worker(rank, world, port, P, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from photo_slam_b200.parallel import GradBuffer, PER_GAUSSIAN, shard_schedule
    params, grads, oracle_c = _rank_gradients(rank, P)
    buf = GradBuffer(P, "cpu")
    assert buf.flat.numel() == P * PER_GAUSSIAN
    for v, g in zip(buf.views(), grads):
        v.copy_(torch.from_numpy(np.ascontiguousarray(g, np.float32)).view_as(v))
    scale = buf.all_reduce()
    assert scale == 1.0 / world
    new = _adam_all(params, [v.numpy() for v in buf.views()], oracle_c, scale)
    sched = shard_schedule(list(range(11)), rank, world)
    q.put((rank, [n.copy() for n in new], sched))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_matches_single_process_mean_gradient():
    P, world = 300, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, new, sched = q.get(timeout=240)
        res[r] = (new, sched)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas identical
    for a, b in zip(res[0][0], res[1][0]):
        assert np.array_equal(a, b)
    # equal to one process applying the mean gradient of both views
    params, g0, oracle_c = _rank_gradients(0, P)
    _, g1, _ = _rank_gradients(1, P)
    summed = [np.asarray(a, np.float32) + np.asarray(b, np.float32) for a, b in zip(g0, g1)]
    single = _adam_all(params, summed, oracle_c, 0.5)
    for a, b in zip(res[0][0], single):
        assert np.allclose(a, b, rtol=1e-6, atol=1e-9)
    # schedule sharding: disjoint, interleaved, K consecutive entries per step, remainder dropped
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7, 9]


def _gather_worker(rank, world, port, P, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from photo_slam_b200.parallel import gather_owned_rows, owner_of_rows
    own = owner_of_rows(P, world)
    truth = [torch.arange(P * k, dtype=torch.float32).view(P, *sh) for k, sh in ((45, (15, 3)), (1, (1,)), (4, (4,)))]
    mine = []
    for t in truth:   # valid on the owned rows, garbage elsewhere (what the p2p step leaves in the moment tensors)
        x = torch.full_like(t, float(-1000 - rank))
        x[own == rank] = t[own == rank]
        mine.append(x)
    gather_owned_rows(mine, rank, world)
    q.put((rank, all(torch.equal(a, b) for a, b in zip(mine, truth))))
    dist.barrier()
    dist.destroy_process_group()


def test_owner_map_and_moment_gather_two_ranks():
    """p2p step host logic: chunk-interleaved ownership and the gather that precedes densification / checkpoints."""
    sys.path.insert(0, ROOT)
    from photo_slam_b200.parallel import CHUNK, owner_of_rows
    own = owner_of_rows(1000, 3)
    assert CHUNK == 128 and own[0] == 0 and own[127] == 0 and own[128] == 1 and own[256] == 2 and own[384] == 0 and own[999] == (999 // 128) % 3
    P, world = 700, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _p2p_worker(rank, world, port, P, q):
    """One rank of the NVLink data-parallel step, emulated on CPU with the host-side protocol specification (parallel.pack_records /
    reduce_records): per-rank gradients from the C oracle -> 80-byte records -> owner-side sum -> Adam on the owned rows only ->
    all-gather of the updated rows (gloo stands in for the peer stores)."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import photo_slam_b200.synthetic as syn
    from photo_slam_b200 import parallel as par
    params, grads, oracle_c = _rank_gradients(rank, P)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
    g = [T(x) for x in grads]
    R, t = syn.random_pose(np.random.default_rng(10 + rank), 0.1, 0.1)
    campos = T(syn.make_camera(96, 64, 80.0, 80.0, R, t)["campos"])
    xyz = T(params[0])
    # the clamp-masked dL/dRGB is what the kernel holds; here it is recovered from the DC gradient (weight C0)
    gm = g[1].reshape(P, 3) / 0.28209479177387814
    # the rank-1 claim the protocol rests on: dL/dsh_k = w_k(dir) * masked dL/dRGB
    w = par.sh_basis_weights(xyz, campos, 3)
    assert torch.allclose(w[:, 1:].unsqueeze(2) * gm.unsqueeze(1), g[2], rtol=2e-4, atol=1e-9)
    epoch = 7
    rec = par.pack_records(g, gm, epoch)
    assert rec.shape == (P, par.REC_FLOATS) and 0 < int((rec[:, 19] == epoch).sum()) < P          # hidden / culled rows send nothing
    # "push": every owner receives every rank's records of its rows; camera centres travel with them
    all_rec = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(all_rec, rec)
    all_cam = [torch.zeros(3) for _ in range(world)]
    dist.all_gather(all_cam, campos)
    own = par.owner_of_rows(P, world) == rank
    summed = par.reduce_records([r[own] for r in all_rec], xyz[own], torch.stack(all_cam), 3, epoch)
    # sharded Adam: this rank updates its rows only (moments exist only here) ...
    new = [T(p).clone() for p in params]
    for i, (p, gsum, lr) in enumerate(zip(params, summed, LRS)):
        rows = T(p)[own]
        pn, _, _ = oracle_c.adam(rows.numpy().ravel(), (gsum.numpy() * np.float32(1.0 / world)).ravel(), np.zeros(rows.numel(), np.float32),
                                 np.zeros(rows.numel(), np.float32), lr, 1)
        new[i][own] = torch.from_numpy(pn).view_as(rows)
    # ... and the all-gather: every replica receives the updated rows from their owners
    par.gather_owned_rows(new, rank, world)
    q.put((rank, [x.numpy().copy() for x in new]))
    dist.barrier()
    dist.destroy_process_group()


def test_p2p_record_protocol_two_ranks_matches_single_process_mean_gradient():
    """The fused NVLink step's protocol (records, owner-side reduction with reconstructed f_rest gradients, sharded Adam, all-gather) on
    CPU with gloo: replicas identical and equal to ONE process applying the mean gradient of both views to every row."""
    P, world = 300, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, world, port, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, new = q.get(timeout=240)
        res[r] = new
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b), "replicas must be bit-identical"
    params, g0, oracle_c = _rank_gradients(0, P)
    _, g1, _ = _rank_gradients(1, P)
    summed = [np.asarray(a, np.float32) + np.asarray(b, np.float32) for a, b in zip(g0, g1)]
    single = _adam_all(params, summed, oracle_c, 0.5)
    for a, b, lr in zip(res[0], single, LRS):
        # the f_rest gradient is reconstructed (w * dL/dRGB) instead of transmitted: agreement in units of one Adam step
        assert np.abs(a.ravel() - b.ravel()).max() <= 0.02 * lr, np.abs(a.ravel() - b.ravel()).max()
