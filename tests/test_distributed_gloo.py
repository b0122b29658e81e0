"""world_size-2 `gloo` test (CPU) of the multi-GPU host protocol of the standardisation.

The product's ``Standardizer.fit`` is: local passes 1 and 2 (about the LOCAL mean) -> this rank's message
``[n, S, C, Q]`` -> ONE all-gather -> merge in rank order (Chan's update to the global mean).  The local passes and the
merge are CUDA kernels; here they are replaced by a numpy TEST DOUBLE with the same message contract (defined in this
file, not in the product) so the protocol - message layout, row counts, the single collective, clip sharding by index -
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
    """Test double: the launch groups in numpy (float64), same message layout and merge formulas as the kernels."""

    def local_stats(self, blocks, noises=None):
        self._blocks = [b.numpy() for b in blocks]
        self._n_local = sum(b.shape[0] for b in self._blocks)

    def local_message(self):
        D = self.n_cols
        x = np.concatenate(self._blocks, axis=0) if self._blocks else np.zeros((0, D))
        n = x.shape[0]
        S = x.sum(axis=0)
        m = S / max(n, 1)
        c = x - m
        self.msg[0] = float(n)
        self.msg[1:1 + D] = torch.from_numpy(S)
        self.msg[1 + D:1 + 2 * D] = torch.from_numpy(c.sum(axis=0))
        self.msg[1 + 2 * D:] = torch.from_numpy((c * c).sum(axis=0))

    def merge(self):
        D = self.n_cols
        M = self.msgs.numpy()
        n = M[:, 0].sum()
        mu = M[:, 1:1 + D].sum(axis=0) / n
        corr = np.zeros(D)
        ssq = np.zeros(D)
        for r in range(M.shape[0]):
            nr = M[r, 0]
            if nr <= 0:
                continue
            d = M[r, 1:1 + D] / nr - mu
            C, Q = M[r, 1 + D:1 + 2 * D], M[r, 1 + 2 * D:]
            corr += C + nr * d
            ssq += Q + 2 * d * C + nr * d * d
        var = (ssq - corr * corr / n) / n
        eps = np.finfo(np.float64).eps
        scale = np.sqrt(var)
        scale[var <= n * eps * var + (n * mu * eps) ** 2] = 1.0          # sklearn _is_constant_feature
        self.mean.copy_(torch.from_numpy(mu)); self.var.copy_(torch.from_numpy(var)); self.scale.copy_(torch.from_numpy(scale))
        self.n_dev[0] = n
        self._n_total = None

    def transform(self, x, out_dtype=torch.float64, out=None, noise=None):
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
        st.fit([mine])                                          # n_total comes from the gathered messages
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
