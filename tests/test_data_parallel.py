"""CPU tests of the multi-GPU host logic with world_size-2 gloo: batch sharding +
the flat gradient arena all-reduce must reproduce single-process gradients on the concatenated batch.
Per-shard gradients come from the oracle (the CUDA kernels cannot run here)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import poet_oracle as O
from poet_b200 import synthetic as S
from poet_b200.data_parallel import FlatGradReducer, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(cfg, P, lo, hi):
    inp = S.make_inputs(cfg)
    g_t, g_R = S.make_cotangents(cfg)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    cap = {}
    O.poet_path_forward(Pr, cfg, [s[lo:hi] for s in inp["srcs"]], [m[lo:hi] for m in inp["masks"]],
                        inp["boxes"][lo:hi], inp["labels"][lo:hi], capture=cap)
    O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t[:, lo:hi], g_R[:, lo:hi]).backward()
    return {k: v.grad for k, v in Pr.items()}


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = S.CONFIGS["tiny16"]
    P = S.make_params(cfg)
    params = [torch.nn.Parameter(v.clone()) for v in P.values()]
    red = FlatGradReducer(params, average=True)
    lo, hi = shard_range(cfg["batch"], rank, world)
    red.zero()
    for p, g in zip(params, _grads(cfg, P, lo, hi).values()):
        if g is not None:                    # transformer.reference_points.* stays zero in the arena
            p.grad += g
    red.all_reduce()
    assert all(p.grad.data_ptr() == red.flat.data_ptr() + off * 4 for p, off in zip(params, red.offsets))
    if rank == 0:
        torch.save({k: p.grad.clone() for k, p in zip(P.keys(), params)}, os.path.join(out_dir, "reduced.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reduced = torch.load(os.path.join(tmp_path, "reduced.pt"))
    cfg = S.CONFIGS["tiny16"]
    full = _grads(cfg, S.make_params(cfg), 0, cfg["batch"])
    for k, g in full.items():
        if g is None:
            assert float(reduced[k].abs().max()) == 0.0, k
            continue
        ref = g / world
        assert float((reduced[k] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-7, k


def test_shard_range_partitions_batch():
    assert [shard_range(16, r, 4) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    try:
        shard_range(10, 0, 4)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_arena_views_are_aligned_and_rebindable():
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
    red = FlatGradReducer(ps)
    assert all(off % 4 == 0 for off in red.offsets)
    ps[1].grad = None
    red.zero()
    assert ps[1].grad is not None and ps[1].grad.data_ptr() == red.flat.data_ptr() + red.offsets[1] * 4
    (ps[0].sum() * 2 + ps[1].sum()).backward()
    assert float(red.flat[:15].sum()) == 30.0 and float(red.flat[16:23].sum()) == 7.0


# ---- segmented (overlapped) all-reduce: host logic -----------------------------------------------------------------
def _seg_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = S.CONFIGS["tiny16"]
    P = S.make_params(cfg)
    names = list(P.keys())
    params = [torch.nn.Parameter(v.clone()) for v in P.values()]
    g = torch.Generator().manual_seed(100 + rank)
    red = FlatGradReducer(params, average=True)
    assert red.plan_overlap(list(zip(names, params)))
    red.zero()
    red.flat.copy_(torch.randn(red.flat.numel(), generator=g))
    mine = red.flat.clone()
    # the markers fire in backward order: decoder first, then the encoder layers from the last to the first
    red.begin_step()
    fired = [("decoder", 0)] + [("encoder", i) for i in reversed(range(cfg["enc_layers"]))]
    for key in fired:
        red.on_marker(key)
    covered = sorted(red._done)
    assert covered and all(a < b for a, b in covered)
    assert all(covered[i][1] <= covered[i + 1][0] for i in range(len(covered) - 1)), "segments overlap"
    assert sum(b - a for a, b in covered) < red.flat.numel(), "something must be left for finish()"
    red.finish()
    done = sorted(red._done)
    assert done[0][0] == 0 and done[-1][1] == red.flat.numel()
    assert all(done[i][1] == done[i + 1][0] for i in range(len(done) - 1)), "every element reduced exactly once"
    torch.save({"mine": mine, "reduced": red.flat.clone()}, os.path.join(out_dir, f"seg{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_segmented_allreduce_covers_the_arena_exactly_once(tmp_path):
    world = 2
    mp.spawn(_seg_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recs = [torch.load(os.path.join(tmp_path, f"seg{r}.pt")) for r in range(world)]
    mean = sum(r["mine"] for r in recs) / world
    for r in recs:
        assert torch.allclose(r["reduced"], mean, rtol=0, atol=1e-6)


def _coarse_worker(rank, world, port, out_dir):
    os.environ["POET_OVERLAP_COARSE"] = "1"
    _seg_worker(rank, world, port, out_dir)


def test_coarse_overlap_plan_reduces_every_element_once(tmp_path):
    """POET_OVERLAP_COARSE: one early collective (the decoder's segment, launched when the last encoder layer's marker
    fires), everything else in finish(): same coverage and same mean as the per-layer plan."""
    world = 2
    mp.spawn(_coarse_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recs = [torch.load(os.path.join(tmp_path, f"seg{r}.pt")) for r in range(world)]
    mean = sum(r["mine"] for r in recs) / world
    for r in recs:
        assert torch.allclose(r["reduced"], mean, rtol=0, atol=1e-6)
