"""Multi-process host logic on CPU (gloo, world_size 2): contiguous sharding + result gather reproduce the unsharded
result exactly.  The per-shard compute here is the plain-C oracle (the GPU path is covered by -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch(eikws):
    from eikws_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 65536, 1048576, 1000003):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, n_total, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import eikws_pkg
    eikws_pkg.load()
    import eikws_b200.synth as synth
    from eikws_b200.sharding import gather_results, shard_range
    from oracle_lib import PortOracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, world, rank)
    clips = synth.synth_clips(hi - lo, first_clip=lo)
    local = torch.from_numpy(PortOracle("l476").run_classifier_i16(clips))
    full = gather_results(local, n_total, world, rank)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo(tmp_path, synth):
    import torch.multiprocessing as mp
    from oracle_lib import PortOracle
    n_total, world = 13, 2  # odd on purpose: ragged shards
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, port, n_total, out), nprocs=world, join=True)
    want = PortOracle("l476").run_classifier_i16(synth.synth_clips(n_total))
    assert np.array_equal(np.load(out), want)
