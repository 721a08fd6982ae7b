"""
World-size-2 `gloo` test of the multi-process host logic (CPU): sample sharding + the one exchange step
(all_reduce of the packed accumulators) + finalisation reproduce the single-process image.  The CPU oracle stands in for
the CUDA renderer here (this file tests the plumbing around it; GPU sharding itself is covered in test_render_gpu.py).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPP, DEPTH, H, W = 6, 3, 24, 32


def _render_shard(rank, world):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    import scenes
    from test_oracle_golden import make_camera
    from diffrp_b200.path_tracing import shard_sample_ids
    cam = make_camera(dict(h=H, w=W), None)
    ids = shard_sample_ids(SPP, rank, world).numpy()
    vao, hs, p, keep = scenes.oracle_inputs(scenes.mixed_scene(24, 12), cam, SPP, DEPTH, seed=4, sample_ids=ids)
    acc, n = oracle.render(oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy()), hs, p)
    return acc, ids


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from diffrp_b200.path_tracing import reduce_accumulators
    acc, ids = _render_shard(rank, world)
    t = reduce_accumulators(torch.from_numpy(acc), world)
    gathered = [None] * world
    dist.all_gather_object(gathered, ids.tolist())
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), t.numpy())
        np.save(os.path.join(out_dir, "ids.npy"), np.array(sorted(sum(gathered, []))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_sharding_and_accumulator_reduce(tmp_path):
    world, port = 2, 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    summed = np.load(tmp_path / "sum.npy")
    assert np.array_equal(np.load(tmp_path / "ids.npy"), np.arange(SPP))  # shards partition the sample set
    whole, _ = _render_shard(0, 1)
    np.testing.assert_allclose(summed, whole, rtol=1e-5, atol=1e-5)
    sys.path.insert(0, ROOT)
    import oracle
    a, b = oracle.finalize(summed, H, W, SPP), oracle.finalize(whole, H, W, SPP)
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-5, atol=1e-5)


def _tile_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from diffrp_b200.path_tracing import gather_tile_accumulators, reduce_accumulators, frame_tiles, tile_rows
    Ht, Wt, tile = 37, 53, 16                      # ragged: partial tiles on both edges, 12 tiles over 2 ranks
    g = torch.Generator().manual_seed(7)
    whole = torch.rand(Ht * Wt, 16, generator=g)
    mine = torch.zeros_like(whole)
    rows = tile_rows(frame_tiles(Ht, Wt, tile)[rank::world], Wt)
    mine[rows] = whole[rows]
    a = gather_tile_accumulators(mine.clone(), Ht, Wt, tile, rank, world)
    b = reduce_accumulators(mine.clone(), world)
    assert torch.equal(a, whole) and torch.equal(b, whole), rank     # gather of disjoint supports == sum, bit for bit
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_gather_equals_allreduce(tmp_path):
    world, port = 2, 31000 + (os.getpid() % 2000)
    mp.spawn(_tile_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0] == 1


def test_sharded_upload_segments_tile_the_packed_buffer():
    """Host logic of the sharded scene upload (flatten.PackedUpload): every rank copies the part of each source run that falls into ITS 1/world
    byte range of the packed buffer; over all ranks every payload byte is copied exactly once, from the right source address."""
    sys.path.insert(0, ROOT)
    from diffrp_b200.flatten import clip_runs, merge_runs, PackedUpload
    rng = np.random.default_rng(0)
    sizes = [int(x) for x in rng.integers(1, 5000, 37)]
    items, off, addr = [], 0, 10_000_000
    for n in sizes:
        items.append((off, addr, n))
        step = -(-n // PackedUpload.ALIGN) * PackedUpload.ALIGN
        off += step
        addr += step if rng.random() < 0.7 else step + 4096          # some sources are laid out like the buffer, some are not
    total = off
    for arena in (None, (10_000_000, addr + 10)):
        runs = merge_runs(items, arena)
        assert sum(n for _, _, n in runs) >= sum(sizes) and (arena is None) == (len(runs) == len(items))
        for world in (1, 2, 3, 8):
            per = -(-total // (world * PackedUpload.ALIGN)) * PackedUpload.ALIGN
            covered = np.zeros(per * world, np.int64)
            source = np.zeros(per * world, np.int64)
            for r in range(world):
                for a, p, n in clip_runs(runs, r * per, (r + 1) * per):
                    assert r * per <= a and a + n <= (r + 1) * per
                    covered[a:a + n] += 1
                    source[a:a + n] = p + np.arange(n)
            for o, p, n in items:                                        # payload bytes: exactly once, from their own source
                assert (covered[o:o + n] == 1).all() and (source[o:o + n] == p + np.arange(n)).all()
            assert covered.max() <= 1


def _reduce_to_rank_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from diffrp_b200.path_tracing import reduce_accumulators
    g = torch.Generator().manual_seed(11)
    parts = [torch.rand(97, 16, generator=g) for _ in range(world)]
    mine = parts[rank].clone()
    out = reduce_accumulators(mine, world, dst=1)                    # options.result_rank = 1: only that rank receives the sum
    ok = True
    if rank == 1:
        ok = torch.allclose(out, sum(parts), rtol=0, atol=1e-6)
    try:                                                             # a shard_world that is not the group's size must fail loudly
        reduce_accumulators(parts[rank].clone(), world + 1)
        ok = False
    except RuntimeError as e:
        ok = ok and "does not match" in str(e)
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(all(flags))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_to_result_rank_and_world_size_check(tmp_path):
    """options.result_rank: `reduce` to one rank instead of the all-reduce; and the guard against a shard_world that does not match the group."""
    world, port = 2, 33000 + (os.getpid() % 2000)
    mp.spawn(_reduce_to_rank_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0] == 1


def test_sharded_exchange_without_a_process_group_raises():
    """A sharded render whose accumulators are never summed would silently return 1/world of the energy: it must raise instead."""
    import pytest
    sys.path.insert(0, ROOT)
    from diffrp_b200.path_tracing import reduce_accumulators, gather_tile_accumulators
    assert not (dist.is_available() and dist.is_initialized())
    acc = torch.zeros(16, 16)
    with pytest.raises(RuntimeError, match="not initialised"):
        reduce_accumulators(acc, 2)
    with pytest.raises(RuntimeError, match="not initialised"):
        gather_tile_accumulators(acc, 4, 4, 2, 0, 2)
    assert reduce_accumulators(acc, 1) is acc and gather_tile_accumulators(acc, 4, 4, 2, 0, 1) is acc   # unsharded: no exchange, no group needed
