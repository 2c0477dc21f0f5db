"""Multi-process (gloo, world_size 2, CPU) test of the N>1 plumbing used by bench.py: frame pairs are
sharded by rank with no data-path collective; only the step time is reduced (MAX) across ranks and the
whole-job value is computed from it.  The compute itself is exercised on the GPU tests; here each rank
stands in with the CPU oracle on its own shard and the shards are checked against the unsharded result."""
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


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # global batch of 4 frame pairs, contiguous batch slices per rank (bench.py: one batch per rank)
    r = np.random.default_rng(0)
    prv = r.standard_normal((4, 6, 7, 4)).astype(np.float32)
    nxt = r.standard_normal((4, 6, 7, 4)).astype(np.float32)
    flo = r.standard_normal((4, 6, 7, 2)).astype(np.float32)
    per = 4 // world
    sl = slice(rank * per, (rank + 1) * per)
    out = oracle.warp_cost_volume(prv[sl], nxt[sl], flo[sl], "tfa", 4)
    # timing reduction exactly as bench.py does it: MAX over ranks, value = units / max time
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros_like(torch.from_numpy(out)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out))          # test-only: to compare with unsharded
    if rank == 0:
        full = oracle.warp_cost_volume(prv, nxt, flo, "tfa", 4)
        q.put((float(t.item()), np.array_equal(torch.cat(gathered).numpy(), full)))
    dist.destroy_process_group()


def test_batch_sharding_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t_max, same = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert abs(t_max - 0.020) < 1e-12          # slowest rank defines the step time
    assert same                                  # shards == unsharded result: no cross-pair term


def test_bench_reference_arm_runs_on_rank0_only(tmp_path):
    import json
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""        # non-zero ranks exit 0 without work


# ------------------------------------------------------------------ row-sharded single frame (config 5)
def _halo_worker(rank, world, port, q, H=41):
    sys.path.insert(0, ROOT)
    import oracle
    from qpwcnet_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = np.random.default_rng(7)
    B, W, C, d = 1, 9, 4, 4
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    flo = (r.standard_normal((B, H, W, 2)) * 1.5).astype(np.float32)
    r0, r1 = sharded.band(H, rank, world)

    def cv_op(p, n, dd, slope):          # the compute is injected: CPU oracle here, libqpwc on the GPU
        return torch.from_numpy(oracle.cost_volume(p.numpy(), n.numpy(), dd, slope))

    def wcv_op(p, n, f, mode, dd, slope):
        return torch.from_numpy(oracle.warp_cost_volume(p.numpy(), n.numpy(), f.numpy(), mode, dd, slope))

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, r0:r1]))
    out = sharded.cost_volume(t(prv), t(nxt), d, 0.1, op=cv_op)
    ok_cv = np.array_equal(out.numpy(), oracle.cost_volume(prv, nxt, d)[:, r0:r1])
    res = {}
    for mode in ("tf", "tfa"):
        o = sharded.warp_cost_volume(t(prv), t(nxt), t(flo), mode, d, 0.1, op=wcv_op)
        ref = oracle.warp_cost_volume(prv, nxt, flo, mode, d)
        # the plain cost volume is bit-identical under sharding; the warp computes its sampling
        # coordinate as float(row) + flow, and a band-local row index rounds differently from the
        # global one (fp32, ~1 ulp of the coordinate) => equal within the cost-volume tolerance
        res[mode] = float(np.abs(o.numpy() - ref[:, r0:r1]).max()) <= 1e-5 * float(np.abs(ref).max())
    q.put((rank, r0, r1, ok_cv, res["tf"], res["tfa"]))
    dist.barrier()
    dist.destroy_process_group()


def _run_halo(world, H=41):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q, H)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rows = 0
    for rank, r0, r1, ok_cv, ok_tf, ok_tfa in got:
        assert ok_cv and ok_tf and ok_tfa, f"rank {rank} rows [{r0},{r1})"
        rows += r1 - r0
    assert rows == H


def test_row_sharded_halo_exchange_world_2():
    _run_halo(2)


def test_row_sharded_halo_exchange_world_3():
    _run_halo(3)          # the middle rank has two neighbours


def test_row_sharded_thin_bands_fall_back_to_all_gather():
    """Bands thinner than the halo (coarse levels of a frame cut into many bands, BASELINE config 5 on
    8 GPUs: 68 rows / 8): the neighbour exchange is replaced by an all-gather of the bands.  World 4:
    H = 41 -> 10-row bands vs the fused op's ~11-row halo; H = 13 -> 3-row bands vs the 4-row halo of
    the plain cost volume."""
    _run_halo(4)
    _run_halo(4, H=13)


# ------------------------------------------------- overlapped, allocation-free level (ShardedLevel)
class _OracleOps:
    """CPU stand-in for qpwcnet_b200.ops with the three entry points ShardedLevel uses."""

    @staticmethod
    def cost_volume_into(out, prv, nxt, d):
        import oracle
        out.copy_(torch.from_numpy(oracle.cost_volume(np.ascontiguousarray(prv.numpy()), np.ascontiguousarray(nxt.numpy()), d)))
        return out

    @staticmethod
    def warp(img, flow, mode):
        import oracle
        return torch.from_numpy(oracle.warp(np.ascontiguousarray(img.numpy()), np.ascontiguousarray(flow.numpy()), mode))


def _level_worker(rank, world, port, q, H):
    sys.path.insert(0, ROOT)
    import oracle
    from qpwcnet_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = np.random.default_rng(11)
    W, C, d, reach = 9, 4, 4, 3
    prv = r.standard_normal((1, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((1, H, W, C)).astype(np.float32)
    flo = np.clip(r.standard_normal((1, H, W, 2)) * 1.5, -reach, reach).astype(np.float32)
    res = []
    # plain cost volume: interior first, halo strips after the exchange -- bit-identical to the unsharded op
    lv = sharded.ShardedLevel(H, W, C, d, device="cpu", ops_module=_OracleOps)
    lv.prv.copy_(torch.from_numpy(prv[:, lv.r0:lv.r1]))
    lv.nxt.copy_(torch.from_numpy(nxt[:, lv.r0:lv.r1]))
    for _ in range(2):                     # twice: buffers are reused, nothing is reallocated
        out = lv.run()
    res.append(bool(np.array_equal(out.numpy(), oracle.cost_volume(prv, nxt, d)[:, lv.r0:lv.r1])))
    # UpFlow pair with a fixed flow-reach budget
    for mode in ("tf", "tfa"):
        lp = sharded.ShardedLevel(H, W, C, d, reach=reach, pair=True, mode=mode, device="cpu", ops_module=_OracleOps)
        lp.prv.copy_(torch.from_numpy(prv[:, lp.r0:lp.r1]))
        lp.nxt.copy_(torch.from_numpy(nxt[:, lp.r0:lp.r1]))
        lp.flow.copy_(torch.from_numpy(flo[:, lp.r0:lp.r1]))
        out = lp.run()
        ref = oracle.warp_cost_volume(prv, nxt, flo, mode, d)
        res.append(float(np.abs(out.numpy() - ref[:, lp.r0:lp.r1]).max()) <= 1e-5 * float(np.abs(ref).max()))
    q.put((rank, lv.r0, lv.r1, res))
    dist.barrier()
    dist.destroy_process_group()


def _run_level(world, H):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_level_worker, args=(r, world, port, q, H)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, r0, r1, res in got:
        assert all(res), f"rank {rank} rows [{r0},{r1}): cost volume / pair tf / pair tfa = {res}"


def test_sharded_level_overlapped_world_2():
    _run_level(2, 24)


def test_sharded_level_overlapped_world_3():
    _run_level(3, 33)      # the middle rank has two neighbours; 11-row bands vs an 8-row halo
