"""Multi-GPU path on real devices: ONE tractogram cut into CSR ranges by
sharding.shard_ranges, one process per GPU over NCCL, kernel 2 writing the 27-double partial block that the single
all-gather moves.  Checks: the df_sl table assembled from the shards is BIT-identical to the single-GPU table
(a polyline's result does not depend on the shard, window or group it lands in), and the bundle means combined in
rank order agree with the single-GPU means to 1e-12.

Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`); skipped on a 1-GPU box.
The gloo/CPU twin of the host logic is tests/test_sharding_cpu.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tractogram(dev):
    """Three bundles, mixed length laws (normal, heavy tail with long polylines, short), a few degenerate polylines."""
    import torch
    from lesion_condition_vae_b200 import synth
    n = torch.cat([synth.torch_lengths("normal", 60_000, 11, dev), synth.torch_lengths("heavy", 30_000, 12, dev),
                   torch.randint(2, 12, (10_000,), device=dev, generator=torch.Generator(device=dev).manual_seed(13))])
    pts, off = synth.torch_random_walk_csr(n, 21, dev)
    pts[int(off[777]) + 1, 1] = float("nan")                       # dropped by the loader filter
    pts[int(off[888]):int(off[889])] = pts[int(off[888])]          # zero length: dropped by the length filter
    bo = np.array([0, 60_000, 90_000, 100_000], dtype=np.int64)
    return pts, off, bo


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from lesion_condition_vae_b200 import _lib, sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = _lib.Context(rank)
        pts, off, bo = _tractogram(dev)                           # the same global tractogram on every rank (seeded)
        S = off.numel() - 1
        bounds = sharding.shard_ranges(off.cpu().numpy(), world)
        shard = sharding.DeviceShard(ctx, pts, off, bounds[rank], bounds[rank + 1], bo)
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            gathered = shard.step(stream.cuda_stream)
        stream.synchronize()
        sums, counts = sharding.combine_partials(gathered.cpu().numpy())
        np.save(os.path.join(tmp, f"out{rank}.npy"), shard.out[:, :shard.S].cpu().numpy())
        np.save(os.path.join(tmp, f"keep{rank}.npy"), shard.keep[:shard.S].cpu().numpy())
        np.save(os.path.join(tmp, f"sums{rank}.npy"), sums)
        np.save(os.path.join(tmp, f"counts{rank}.npy"), counts)
        if rank == 0:                                             # the single-GPU answer for the whole tractogram
            whole = sharding.DeviceShard(ctx, pts, off, 0, S, bo)
            whole.compute(0)
            ctx.synchronize()
            np.save(os.path.join(tmp, "out_single.npy"), whole.out.cpu().numpy())
            np.save(os.path.join(tmp, "keep_single.npy"), whole.keep.cpu().numpy())
            np.save(os.path.join(tmp, "partial_single.npy"), whole.partial.cpu().numpy())
            np.save(os.path.join(tmp, "bounds.npy"), bounds)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_tractogram_equals_single_gpu(tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    from lesion_condition_vae_b200 import sharding
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ld = lambda name: np.load(tmp_path / name)
    bounds = ld("bounds.npy")
    assert bounds[0] == 0 and np.all(np.diff(bounds) > 0)
    out = np.concatenate([ld(f"out{r}.npy") for r in range(world)], axis=1)
    keep = np.concatenate([ld(f"keep{r}.npy") for r in range(world)])
    single, keep1 = ld("out_single.npy"), ld("keep_single.npy")
    assert out.shape == single.shape
    assert np.array_equal(keep, keep1) and (keep != 3).sum() >= 2                       # the degenerate polylines are in
    assert np.array_equal(out.view(np.uint64), single.view(np.uint64)), "sharded df_sl table is not bit-identical"
    # every rank combined the same partials in the same order
    for r in range(1, world):
        assert np.array_equal(ld(f"sums{r}.npy"), ld("sums0.npy")) and np.array_equal(ld(f"counts{r}.npy"), ld("counts0.npy"))
    s1, c1 = sharding.combine_partials(ld("partial_single.npy")[None])
    n1, m1 = sharding.means_from_partials(s1, c1)
    nN, mN = sharding.means_from_partials(ld("sums0.npy"), ld("counts0.npy"))
    assert np.array_equal(n1, nN) and np.array_equal(c1, ld("counts0.npy"))           # counts bit-exact
    fin = np.isfinite(m1)
    assert np.array_equal(fin, np.isfinite(mN)) and np.array_equal(m1[~fin], mN[~fin], equal_nan=True)
    assert np.all(np.abs(mN[fin] - m1[fin]) <= 1e-12 * np.abs(m1[fin]) + 1e-15)
