"""world_size-2 `gloo` test (CPU) of the multi-GPU host protocol of the standardisation.

The product's ``Standardizer.fit`` is: local pass 1 -> all-reduce [sum x, n] -> local pass 2 ->
all-reduce [sum (x-mean), sum (x-mean)^2] -> finish.  The local passes are CUDA kernels; here they are
replaced by a numpy TEST DOUBLE with the same accumulator contract (defined in this file, not in the
product) so the protocol - accumulator layout, row counts, the two collectives, clip sharding by index -
runs over a real 2-rank process group and is compared with the oracle on the full dataset.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asr_b200 import sharding
from asr_b200.frontend import Standardizer
from oracle import cmvn_ref as cr


class _NumpyLocalPasses(Standardizer):
    """Test double: the three local launch groups in numpy (float64), same accumulator layout as the kernels."""

    def pass1_local(self, blocks):
        D = self.n_cols
        self.acc1.zero_()
        rows = 0
        for x in blocks:
            self.acc1[:D] += torch.from_numpy(x.numpy().sum(axis=0))
            rows += x.shape[0]
        self.acc1[D] = float(rows)

    def pass2_local(self, blocks, n_total):
        D = self.n_cols
        self.n_total = int(n_total)
        self.mean.copy_(self.acc1[:D] / self.n_total)
        self.acc2.zero_()
        m = self.mean.numpy()
        for x in blocks:
            c = x.numpy() - m
            self.acc2[:D] += torch.from_numpy(c.sum(axis=0))
            self.acc2[D:] += torch.from_numpy((c * c).sum(axis=0))

    def finish(self):
        D, n = self.n_cols, self.n_total
        var = self.acc2[D:] / n - (self.acc2[:D] / n) ** 2
        self.var.copy_(var)
        scale = torch.sqrt(var)
        eps = np.finfo(np.float64).eps
        constant = var <= n * eps * var + (n * self.mean * eps) ** 2      # sklearn _is_constant_feature
        scale[constant] = 1.0
        self.scale.copy_(scale)

    def transform(self, x, out_dtype=torch.float64, out=None):
        return ((x - self.mean) / self.scale).to(out_dtype)


def _worker(rank, world, port, n_rows, n_cols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        full = rng.standard_normal((n_rows, n_cols)) * 4 + 3
        full[:, 2] = 0.0                                        # constant column -> scale 1 (zero-padded frames)
        lo, hi = sharding.shard_bounds(n_rows, rank, world)
        mine = torch.from_numpy(full[lo:hi].copy())
        st = _NumpyLocalPasses(n_cols, device="cpu", distributed=True)
        st.fit([mine])                                          # n_total comes from the all-reduced count
        out = st.transform(mine)
        q.put((rank, lo, hi, st.n_total, st.mean.numpy().copy(), st.var.numpy().copy(), st.scale.numpy().copy(),
               out.numpy().copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_rows", [101, 64])
def test_two_rank_standardisation_equals_single_process(n_rows):
    world, n_cols = 2, 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, n_cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(11)
    full = rng.standard_normal((n_rows, n_cols)) * 4 + 3
    full[:, 2] = 0.0
    mean, var, scale = cr.column_stats(full)
    want = cr.standardize_dataset(full[:1], full[1:2], full[2:])
    want = np.concatenate(want)
    rows = []
    for rank, lo, hi, n_total, m, v, s, out in res:
        assert n_total == n_rows
        np.testing.assert_allclose(m, mean, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(v, var, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(s, scale, rtol=1e-10, atol=1e-12)
        rows.append(out)
    # concatenating the ranks' rows in rank order reproduces the single-process row order
    np.testing.assert_allclose(np.concatenate(rows), want, rtol=1e-9, atol=1e-9)
