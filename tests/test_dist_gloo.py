"""N > 1 host path on CPU: two gloo ranks each render their tile partition (with the CPU oracle standing
in for the device renderer), sum-reduce the framebuffer with adapt_b200.dist.reduce_framebuffer, and
rank 0 must hold exactly the single-rank image (partition invariance + gather-by-sum)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      ADAPT_QUIET="1")
    sys.path.insert(0, ROOT)
    import torch
    from adapt_b200._lib import pack_scene
    from adapt_b200.dist import init_process_group, reduce_framebuffer, tile_partition
    from adapt_b200.parsers.xml_parser import scene_parsing
    from adapt_b200.scenes import DEFAULT_ROOT
    from oracle.pt_oracle import OracleScene
    r, _lr, w = init_process_group("gloo")
    assert (r, w) == (rank, world)
    e, a, o, c = scene_parsing(os.path.join(DEFAULT_ROOT, "csphere"), "balls-mono.xml")
    c["film"]["width"] = 48; c["film"]["height"] = 40
    osc = OracleScene(pack_scene(e, a, o, c, seed=11))
    mine = tile_partition(48, 40, rank, world, tile=8)
    acc, cn = osc.render(3, pixel_list=mine, n_threads=2)
    # non-owned pixels stay exactly zero
    mask = np.zeros(48 * 40, bool); mask[mine] = True
    assert not acc.reshape(-1, 3)[~mask].any()
    buf = torch.from_numpy(acc)
    reduce_framebuffer(buf, dst=0)
    if rank == 0:
        full, _ = osc.render(3, n_threads=2)
        np.save(out_path, np.stack([buf.numpy(), full]))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_reduce_equals_single_rank(tmp_path, scene_root, oracle_lib):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "fb.npy")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    got, full = np.load(out)
    np.testing.assert_array_equal(got, full)
