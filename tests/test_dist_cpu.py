"""CPU: the N>1 path (batch sharding + one weight broadcast, no per-step collective) with gloo, world_size 2."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from diffute_b200 import arch, dist as dd, synthetic
    r, w, _ = dd.init_from_env("gloo")
    shapes = dict(list(arch.vae_param_shapes().items())[:12])
    sd = synthetic.make_state_dict(shapes) if r == 0 else None
    got = dd.broadcast_state_dict(sd, shapes, "cpu")
    ref = synthetic.make_state_dict(shapes)
    same = all(torch.equal(got[k], ref[k]) for k in shapes)
    lo, hi = dd.shard_range(5, r, w)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 3)
    counts = [dd.shard_range(5, i, w)[1] - dd.shard_range(5, i, w)[0] for i in range(w)]
    allimgs = dd.gather_images(local, counts)
    mx = dd.max_over_ranks(float(r + 1), "cpu")
    dd.barrier()
    q.put((r, same, (lo, hi), None if allimgs is None else allimgs[:, 0].tolist(), mx))
    dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                       # identical weights on both ranks
    assert res[0][2] == (0, 3) and res[1][2] == (3, 5)  # contiguous, earlier ranks larger
    assert res[0][3] == [0.0, 1.0, 2.0, 3.0, 4.0] and res[1][3] is None
    assert res[0][4] == res[1][4] == 2.0


def test_shard_range_partitions():
    from diffute_b200.dist import shard_range
    for total in (0, 1, 7, 8, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
