"""CPU: synthetic generator identity (numpy vs torch) and the N>1 host logic over gloo."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle
from imagecodecs_b200.sharding import shard_range
from imagecodecs_b200.synth import synth_batch as torch_batch


@pytest.mark.parametrize("kind", ["photo", "noise"])
@pytest.mark.parametrize("c", [1, 3, 4])
def test_torch_generator_equals_numpy_generator(kind, c):
    a = oracle.synth_batch(3, 37, 21, c, kind, first=5)
    b = torch_batch(3, 37, 21, c, kind, first=5).numpy()
    assert np.array_equal(a, b)


def test_generator_matches_survey_definition():
    # hand-evaluated: photo, n=0, x=y=c=0, seed=1 -> low nibble of mix(GOLD)
    z = 0x9E3779B97F4A7C15
    M = (1 << 64) - 1
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z ^= z >> 31
    assert oracle.synth_image(4, 4, 3)[0, 0, 0] == (z & 15)
    assert oracle.synth_image(4, 4, 3, kind="noise")[0, 0, 0] == (z & 255)
    assert (oracle.synth_image(4, 4, 4)[..., 3] == 255).all()


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 256, 16384):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_images, ret):
    import torch.distributed as dist
    from imagecodecs_b200.sharding import max_over_ranks, shard_range
    from tests.emu.emu import emu_encode
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_images, world, rank)
    # each rank encodes ITS shard only (emulated kernel stands in for the GPU here)
    batch = oracle.synth_batch(hi - lo, 40, 24, 3, "photo", first=lo)
    scans, sizes, status = emu_encode(batch, 0, 2, 0, n_ctas=2)
    # the only communication of the multi-GPU path: the timing reduction
    t = max_over_ranks(10.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, [bytes(s) for s in scans], t))
    if rank == 0:
        ret.put(gathered)
    dist.destroy_process_group()


def test_two_rank_gloo_shard_invariance():
    """world_size 2: outputs of image i do not depend on which rank / shard encoded it."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    n_images, world = 5, 2
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs: p.start()
    gathered = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert [g[3] for g in gathered] == [11.0, 11.0]                  # max over ranks
    per_image = {}
    for lo, hi, scans, _ in gathered:
        for i, s in zip(range(lo, hi), scans):
            per_image[i] = s
    assert sorted(per_image) == list(range(n_images))
    whole = oracle.synth_batch(n_images, 40, 24, 3, "photo")
    for i in range(n_images):
        assert oracle.oracle_headers(40, 24, 3, 0, 0, 2) + per_image[i] == oracle.oracle_encode(whole[i], 0, 2, 0)
