"""Spatially sharded map (BASELINE config 5 / SURVEY 8e): x-slabs of cube columns per rank with a
1 m halo, queries answered by the owning rank, 32 sums allreduced per evaluation."""
import os
import subprocess
import sys

import numpy as np
import pytest

import harness
import oracle
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_shard_filters_partition_the_work_single_gpu(s2m, built):
    """Two 'ranks' of a 2-way sharding as two contexts on one GPU, no communicator: the sums each
    rank forms at the first evaluation add up to the unsharded sums, and each holds part of the map."""
    truth, odom, frames = harness.sequence(20261018, "HDL64", 10)
    O = oracle.Oracle(0.4, 0.8)
    for f in range(8):
        O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    cm, sm = O.get_map(0), O.get_map(1)
    full = s2m.Registrar(0.4, 0.8, trace=True)
    parts = [s2m.Registrar(0.4, 0.8, trace=True, shard_rank=r, shard_world=2) for r in range(2)]
    for R in [full] + parts:
        R.map_upload(cm, sm)
    assert len(full.map_download(1)) == len(sm)
    sizes = [len(R.map_download(0)) + len(R.map_download(1)) for R in parts]
    assert all(0 < n < len(cm) + len(sm) for n in sizes) and sum(sizes) >= len(cm) + len(sm)
    c, s = frames[8]
    for R in [full] + parts:
        R.register(c, s, odom[8, :4], odom[8, 4:])
    _, sums_full, _, _, _ = full.trace_lm(0)
    sums = sum(R.trace_lm(0)[1] for R in parts)
    assert np.allclose(sums, sums_full, rtol=1e-10, atol=1e-10 * np.abs(sums_full).max())
    ne = sum(R.stats.n_edge[0] for R in parts)
    npl = sum(R.stats.n_plane[0] for R in parts)
    assert (ne, npl) == (full.stats.n_edge[0], full.stats.n_plane[0])
    # ownership: every gated query was answered by exactly one rank
    for cls in (0, 1):
        u_full = full.trace_knn(0, cls)[2]
        u = sum(R.trace_knn(0, cls)[2].astype(int) for R in parts)
        assert np.array_equal(u, u_full.astype(int))


def test_sharded_registration_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED_OK" in out.stdout
