"""Host logic of the multi-GPU acquisition pass on CPU: shard ranges, the 16-byte record,
the (max value, min index) merge, and the all-gather over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distbo_b200 import sharding as S


@pytest.mark.parametrize("M,world", [(1 << 20, 8), (300000, 8), (1000, 3), (5, 4), (128, 2), (129, 2), (0, 2)])
def test_shard_ranges_partition_the_candidates(M, world):
    edges = [S.shard_range(M, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == M
    for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
        assert b == c and a <= b
    for lo, hi in edges:
        assert lo % 128 == 0 or lo == M          # shards start on a scoring-tile boundary
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 128 + 127


def test_merge_is_numpy_argmax():
    rs = np.random.RandomState(0)
    for _ in range(200):
        ys = rs.randint(0, 5, size=64).astype(np.float64)        # many ties
        cut = sorted(rs.choice(np.arange(1, 64), size=3, replace=False))
        parts = np.split(np.arange(64), cut)
        vals = [ys[p].max() for p in parts]
        idxs = [p[np.argmax(ys[p])] for p in parts]
        v, i = S.merge_best(vals, idxs)
        assert i == int(np.argmax(ys)) and v == ys.max()


def test_merge_ignores_empty_shards_and_nan():
    v, i = S.merge_best([S.NEG_INF, 0.0, float("nan")], [S.IDX_SENTINEL, 7, 3])
    assert (v, i) == (0.0, 7)


def test_record_round_trip_is_bit_exact():
    for val in [0.0, -0.0, 5e-324, 1.7976931348623157e308, -3.25, float("-inf")]:
        rec = S.pack_best(torch.tensor([val], dtype=torch.float64), torch.tensor([2 ** 40 + 3]))
        vals, idxs = S.unpack_best(rec)
        assert np.array([val]).tobytes() == vals.numpy().tobytes()
        assert int(idxs[0]) == 2 ** 40 + 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ys, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = S.ShardGroup()
        lo, hi = g.range(len(ys))
        part = ys[lo:hi]
        if len(part):
            j = int(np.argmax(part))
            val = torch.tensor([part[j]], dtype=torch.float64)
            idx = torch.tensor([lo + j], dtype=torch.int64)
        else:
            val = torch.tensor([S.NEG_INF], dtype=torch.float64)
            idx = torch.tensor([S.IDX_SENTINEL], dtype=torch.int64)
        state = torch.full((4,), float(rank))
        g.broadcast_fit([state], src=0)
        q.put((rank, g.argmax(val, idx), state.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [1000, 100])
def test_two_rank_gloo_argmax_equals_single_process(M):
    rs = np.random.RandomState(M)
    ys = np.round(rs.randn(M), 1)                                  # ties across the shard boundary
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ys, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, (v, i), state in out:
        assert i == int(np.argmax(ys)) and v == float(ys.max())
        assert state == [0.0] * 4                                  # broadcast from rank 0 arrived


# ---------------------------------------------------------------------------------------------------------------
# The fit -> score hand-off INSIDE the product API (gp_regressor.maximise_likhd / _prepare_predict, reference call
# site bayes_opt/bayesian_optimization.py:287-305): with a process group the fitting rank alone runs the Adam loop
# and computes the posterior state; the other ranks receive the loss and the broadcast state.  The protocol is host
# logic, exercised here over gloo with CPU tensors standing in for the engine's device buffers.
class _StubParams:
    def __init__(self):
        self.flat = torch.zeros(5, dtype=torch.float64)
        self.step = 0


class _StubEngine:
    def __init__(self, N=6, T=2):
        self.dev = torch.device("cpu")
        self.N, self.T, self.P = N, T, 3
        z = lambda *s: torch.zeros(*s, dtype=torch.float64)
        self.W, self.V, self.theta_s, self.sqnorm, self.kvec = z(8, 8), z(8), z(8, 2), z(8), z(8)
        self.params = _StubParams()
        self.W_valid = False

    def close(self):
        pass


def _handoff_worker(rank, world, port, fail, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from distbo_b200 import _lib
        from distbo_b200.train.gp_regressor import gp_regressor
        g = gp_regressor("dist")
        g.engine, g.algorithm = _StubEngine(), "GP"
        trained = []

        def fake_train(max_epochs, first_early_stop_epoch):
            trained.append(rank)
            g.last_epoch = 41
            if fail == "fit":
                raise _lib.DistboError("Cholesky decomposition was not successful")
            return -12.5 - rank

        def fake_posterior():
            if fail == "posterior":
                raise _lib.DistboError("Cholesky decomposition was not successful")
            e = g.engine
            for t in (e.W, e.V, e.theta_s, e.sqnorm, e.kvec, e.params.flat):
                t.fill_(7.0 + rank)
            g._KE_all = torch.full((e.T + 1, e.T + 1), 3.0 + rank, dtype=torch.float64)

        g._train, g._compute_posterior_state = fake_train, fake_posterior
        out = {"rank": rank}
        try:
            # the tail of maximise_likhd (everything before it needs the CUDA engine)
            if g.shards.rank != g.fit_rank:
                out["loss"] = g._receive_fit_result()
            else:
                loss = failure = None
                try:
                    loss = g._train(10, 5)
                except _lib.DistboError as exc:
                    failure = exc
                g._send_fit_result(loss, failure)
                if failure is not None:
                    raise failure
                out["loss"] = loss
            g._fit_generation += 1
            g._prepare_predict()
            e = g.engine
            out["state"] = [float(t.reshape(-1)[0]) for t in (e.W, e.V, e.theta_s, e.sqnorm, e.kvec, e.params.flat,
                                                                g._KE_all)]
            out["last_epoch"], out["W_valid"] = g.last_epoch, bool(e.W_valid)
            g._prepare_predict()                       # same generation: no second collective (would hang if it did)
        except _lib.DistboError as exc:
            out["error"] = str(exc)
        out["trained"] = trained
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail", [None, "fit", "posterior"])
def test_two_rank_fit_handoff_protocol(fail):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_handoff_worker, args=(r, 2, port, fail, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in procs), key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0]["trained"] == [0] and out[1]["trained"] == []      # only the fitting rank runs the Adam loop
    if fail is None:
        for o in out:
            assert o["loss"] == -12.5 and o["last_epoch"] == 41      # rank 0's result on every rank
            assert o["state"] == [7.0] * 6 + [3.0] and o["W_valid"]  # rank 0's posterior state on every rank
    else:
        for o in out:                                                # a failed factorisation raises on EVERY rank
            assert "not successful" in o["error"]
