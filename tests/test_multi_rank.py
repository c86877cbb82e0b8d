"""The N>1 logic on CPU: world_size-2 gloo processes, each rendering its share of the samples (with the CPU restatement
standing in for the GPU renderer — test infrastructure), one reduce to rank 0; the result equals the single-process frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support as S

b2pt = S.b2pt


def test_sample_range_partitions():
    from b2pt.multigpu import sample_range
    for world in (1, 2, 3, 4, 8):
        for spp in (0, 1, 7, 8, 2048, 2051):
            parts = [sample_range(r, world, spp) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == spp
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sample_range(2, 2, 8)


def _worker(rank, world, port, spp_total, out_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import scenes
    import support as S2
    from b2pt.multigpu import reduce_frame, sample_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, _ = scenes.two_triangle_scene()
    pto = S2.Restated(sc)
    begin, count = sample_range(rank, world, spp_total)
    fb = torch.from_numpy(pto.render_frame(begin, count, spp_total))
    reduce_frame(fb, dist, 0)
    if rank == 0:
        np.save(out_path, fb.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_compose_to_the_single_process_frame(tmp_path):
    import scenes
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    spp = 6
    out = str(tmp_path / "fb.npy")
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    got = np.load(out)
    sc, _ = scenes.two_triangle_scene()
    want = S.Restated(sc).render_frame(0, spp, spp)
    assert got.shape == want.shape and want.mean() > 1e-3
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7)  # same samples, summed in a different order
