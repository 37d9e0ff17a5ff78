"""N>1 host logic on CPU: world_size-2 (and 3, ragged) gloo processes shard the image axis, run the path's CPU
restatement on their shard with their slice of the global noise, gather the per-image rows, and must reproduce the
single-process result exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from humaniflow_b200.sharding import gather_rows, sample_diversity_rows, shard, shard_range


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(B, N):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from util import make_model, smpl_data
    m, sd, cfg = make_model(18, seed=0)
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(B, 512, generator=g).abs()
    z = torch.randn(B, N, 23, 3, generator=g) * 0.6
    se = torch.randn(B, N, 10, generator=g)
    return sd, cfg, smpl_data(real_regs=False), feats, z, se


def _rows(sd, cfg, data, feats, z, se):
    from oracle import model as om
    from oracle import smpl as osmpl
    from humaniflow_b200.synthetic import SMPL_PARENTS
    B, N = z.shape[:2]
    out = om.forward(sd, cfg, SMPL_PARENTS, input_feats=feats, compute_point_est=False, num_samples=N, shape_eps=se, base_noise=z)
    R = out['pose_rotmats_samples'].reshape(B * N, 23, 3, 3)
    glob = out['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3)
    _, joints = osmpl.smpl_forward(data, out['shape_samples'].reshape(B * N, 10), R, glob, pose2rot=False)
    return sample_diversity_rows(joints, B, N)


def _worker(rank, world, port, B, N, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd, cfg, data, feats, z, se = _problem(B, N)
    local = _rows(sd, cfg, data, shard(feats, world, rank), shard(z, world, rank), shard(se, world, rank))
    allrows = gather_rows(local, num_images=B)
    pending = gather_rows(local, num_images=B, async_op=True)      # the handle form bench.py uses
    assert torch.equal(pending.result(), allrows)
    if rank == 0:
        q.put(allrows.clone())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,B', [(2, 4), (3, 5)])
def test_sharded_equals_single_process(world, B):
    N = 3
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.set_num_threads(2)
    ref = _rows(*_problem(B, N))
    assert got.shape == (B, 1)
    assert torch.allclose(got, ref, atol=1e-6, rtol=1e-6)


def test_shard_range_partitions():
    for n in (1, 5, 32, 33):
        for w in (1, 2, 3, 8):
            rs = [shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
