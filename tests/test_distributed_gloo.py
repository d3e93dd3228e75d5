"""CPU, world_size 2, gloo: the N>1 host logic — disjoint game sharding, distinct per-rank seeds and the
per-generation sample gather (the only collective on the path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from galvanise_zero_b200 import distributed as gzdist


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    srec = 40
    n_local = 5 + 3 * rank                      # ragged on purpose; rank 1 contributes more
    recs = torch.full((n_local, srec), rank + 1, dtype=torch.int32)
    recs[:, 3] = torch.arange(n_local, dtype=torch.int32)
    got = gzdist.gather_sample_records(recs, dst=0)
    empty = gzdist.gather_sample_records(recs[:0], dst=0)
    if rank == 0:
        assert got.shape == (5 + 8, srec)
        assert (got[:5, 0] == 1).all() and (got[5:, 0] == 2).all()
        assert got[5:, 3].tolist() == list(range(8))
        assert empty.shape[0] == 0
        np.save(os.path.join(out_dir, 'ok.npy'), np.array([1]))
    else:
        assert got is None and empty is None
    dist.destroy_process_group()


def test_gather_sample_records_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), 'ok.npy'))


def test_shards_are_disjoint_and_cover():
    for world in (1, 2, 4, 8):
        for total in (4096, 4097, 7):
            spans = [gzdist.shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert len({gzdist.rank_seed(0xB200, r) for r in range(8)}) == 8
