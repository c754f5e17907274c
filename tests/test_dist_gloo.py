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
