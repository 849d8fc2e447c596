"""CPU, world_size=2 over gloo: the sharding logic of the multi-GPU path.

The frustum path shards by independent units: every rank processes its own
B frustums with replicated weights and there is no data-path collective; the only exchange is the
final all-gather of the per-rank result block (bench.py).  Here two CPU ranks run the ORACLE on
their shard of a batch and all-gather the result; the gathered tensor must equal the single-process
run on the whole batch — the property the N>1 bench relies on."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frustum_convnet_b200 import config, synth
    from frustum_convnet_b200.sharding import shard_slice, pack_outputs
    from oracle import model as om
    torch.set_num_threads(1)
    cfg, w = config.load_workload("car")
    sd = om.to_torch_state(synth.make_state_dict(w["arch"], 3, "KITTI", seed=13))
    B = 4
    data = synth.make_frustums("car", B, seed=31, max_depth=8.75)   # T = (35, 18, 9, 5)
    sl = shard_slice(B, rank, world)
    mine = {k: v[sl] for k, v in data.items()}
    out = om.pointnet_det_eval(mine, sd, cfg.DATA.HEIGHT_HALF, w["arch"].nsample,
                               config.DATASET_INFO["KITTI"].MEAN_SIZE_ARRAY)
    flat = pack_outputs(out)
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        full = om.pointnet_det_eval(data, sd, cfg.DATA.HEIGHT_HALF, w["arch"].nsample,
                                    config.DATASET_INFO["KITTI"].MEAN_SIZE_ARRAY)
        ref = torch.cat([pack_outputs([o[shard_slice(B, r, world)] for o in full]) for r in range(world)])
        ret["ok"] = bool(torch.equal(torch.cat(gathered), ref))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_process():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True


def test_shard_slice_partitions_the_batch():
    from frustum_convnet_b200.sharding import shard_slice
    for B in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            idx = []
            for r in range(world):
                sl = shard_slice(B, r, world)
                idx += list(range(B))[sl]
            assert idx == list(range(B))
