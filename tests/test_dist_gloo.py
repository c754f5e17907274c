"""world_size-2 gloo test (CPU) of the N>1 path: scene sharding + the reductions bench.py uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (path setup via conftest)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests", "golden")]
    import vlsat_b200 as V
    from vlsat_b200 import dist as vd, synth
    from oracle import vlsat_oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = synth.make_real_shaped_batch(5, seed=3, points_per_object=16)
        mine = vd.shard_for_rank(full)
        # every scene lands on exactly one rank
        n_nodes = vd.sum_over_ranks(mine.obj_points.shape[0])
        n_edges = vd.sum_over_ranks(mine.edge_indices.shape[1])
        n_scenes = vd.sum_over_ranks(mine.num_scenes)
        # scene-local stage (graph attention layer) on the shard == the same rows of the full-batch result
        layer = V.GraphEdgeAttenNetwork(4, 64, 32, 32, DROP_OUT_ATTEN=0.5)
        import cases as C
        sd = C.seeded_state(layer, 1)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(full.obj_points.shape[0], 64, generator=g)
        e = torch.randn(full.edge_indices.shape[1], 32, generator=g)
        fx, fe, _ = O.gat_layer(sd, "", x, e, full.edge_indices, 4)
        keep_n = torch.isin(full.batch_ids.view(-1), torch.arange(rank, 5, world))
        keep_e = keep_n[full.edge_indices[0]]
        sx, se, _ = O.gat_layer(sd, "", x[keep_n], e[keep_e], mine.edge_indices, 4)
        ok = torch.allclose(sx, fx[keep_n], atol=1e-6) and torch.allclose(se, fe[keep_e], atol=1e-6)
        slowest = vd.max_over_ranks(1.0 + rank)          # the bench's max-over-ranks timing reduction
        q.put((rank, n_nodes, n_edges, n_scenes, bool(ok), slowest, full.obj_points.shape[0], full.edge_indices.shape[1]))
    finally:
        dist.destroy_process_group()


def test_scene_sharding_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n_nodes, n_edges, n_scenes, ok, slowest, tot_n, tot_e in res:
        assert n_nodes == tot_n and n_edges == tot_e and n_scenes == 5
        assert ok, f"rank {rank}: shard result differs from the full-batch rows"
        assert slowest == 2.0


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests", "golden")]
    import vlsat_b200  # noqa: F401
    from vlsat_b200 import dist as vd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                               # same initial weights on both ranks
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
        for p in net[3].parameters():
            p.requires_grad_(False)                        # frozen module (clip_adapter)
        unused = torch.nn.Linear(4, 4)                     # trainable but never used (triplet_projector_3d): grad stays None
        params = list(net.parameters()) + list(unused.parameters())
        opt = torch.optim.SGD([p for p in params if p.requires_grad], lr=0.5)
        red = vd.GradientAllReducer(params, chunk_elems=64)            # flat buffer, several chunks per tensor
        red.attach(opt)
        g = torch.Generator().manual_seed(100 + rank)      # different data per rank
        x = torch.randn(6, 8, generator=g)
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in params if p.grad is not None]
        before = [p.detach().clone() for p in params]
        opt.step()                                         # hook averages, then SGD applies
        q.put((rank, [t.numpy() for t in local], [(b - p.detach()).numpy() / 0.5 for b, p in zip(before, params) if p.requires_grad and p.grad is not None],
               red.last_bytes, unused.weight.grad is None))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo():
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, a0, nb0, un0), (_, l1, a1, nb1, un1) = res
    assert un0 and un1 and nb0 == nb1 == sum(x.size for x in l0) * 4
    for g0, g1, s0, s1 in zip(l0, l1, a0, a1):
        mean = (g0 + g1) / 2
        assert np.allclose(s0, mean, atol=1e-6) and np.allclose(s1, mean, atol=1e-6)    # both ranks applied the mean gradient
