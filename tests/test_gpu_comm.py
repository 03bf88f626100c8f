"""Multi-GPU inside the library (pn_comm_*, pn_sharded_query_knn_dev, pn_tree_replicate): NCCL is loaded, the chunked
scan -> exchange -> merge pipeline returns exactly the unsharded answer, for both exchange modes.  With one GPU the
communicator has a single rank (every collective still runs); with two or more, all ranks live in this process
(pn_comm_create_all), one host thread each."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import petal_neighbors_b200 as pn
    return pn


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("dtype,n,d,nq,k", [(np.float32, 30000, 16, 160000, 10), (np.float64, 20000, 3, 5000, 7), (np.float32, 500, 64, 300, 10)])
def test_single_rank_pipeline(pn, oracle, dtype, n, d, nq, k):
    import torch
    from petal_neighbors_b200 import parallel, synth
    pts = synth.uniform(n, d, 71, dtype)
    Q = synth.uniform(nq, d, 72, dtype)
    comm = parallel.Comm(parallel.Comm.unique_id(), 1, 0, 0)
    st = parallel.ShardedBallTree(pts, comm)
    qd = torch.from_numpy(Q).cuda()
    sample = np.arange(0, nq, max(1, nq // 300))[:300]
    oi, od = oracle.brute_knn(pts, Q[sample], k)
    for mode in (parallel.PN_EXCHANGE_ALLGATHER, parallel.PN_EXCHANGE_SLICE):
        gi, gd = st.query_batch_dev(qd, k, exchange=mode)
        gi, gd = gi.cpu().numpy(), gd.cpu().numpy()
        assert gi.shape == (nq, k)
        assert np.array_equal(gi[sample].astype(np.uint64), oi.astype(np.uint64)) and np.array_equal(bits(gd[sample]), bits(od))
        s = st.stats
        assert s["rows_out"] == nq and s["nccl_calls"] >= s["n_chunks"] >= 1 and s["nccl_bytes_sent"] == 0
        if nq >= 160000:
            assert s["n_chunks"] >= 2
    t2 = parallel.replicate(st.tree, comm, 0)
    assert t2 is st.tree
    comm.close()


def _ranks_in_threads(fn, world):
    out, err = [None] * world, [None] * world

    def run(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa
            err[r] = e

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    for e in err:
        if e is not None:
            raise e
    return out


def test_two_or_more_gpus_in_one_process(pn, oracle):
    import torch
    from petal_neighbors_b200 import parallel, synth
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least two GPUs")
    world = 1 << (min(ng, 8).bit_length() - 1)
    n, d, nq, k = 200000, 32, 170000, 10
    pts = synth.fast_gaussian_mixture(n, d, 9, n_centers=256, sigma=0.1, clip=True)
    Q = synth.fast_gaussian_mixture(nq, d, 10, n_centers=256, sigma=0.1, clip=True)
    comms = parallel.Comm.create_all(list(range(world)))
    sample = np.arange(0, nq, nq // 400)[:400]
    oi, od = oracle.brute_knn(pts, Q[sample], k)

    def rank(r):
        torch.cuda.set_device(r)
        st = parallel.ShardedBallTree(pts, comms[r])
        qd = torch.from_numpy(Q).cuda(r)
        res = {}
        for mode in (parallel.PN_EXCHANGE_ALLGATHER, parallel.PN_EXCHANGE_SLICE):
            gi, gd = st.query_batch_dev(qd, k, exchange=mode)
            res[mode] = (gi.cpu().numpy(), gd.cpu().numpy(), dict(st.stats))
        # replication of rank 0's (full) tree
        full = pn.BallTree.euclidean(pts, device=0) if r == 0 else None
        rep = parallel.replicate(full, comms[r], 0)
        ri, rd = rep.query_batch(Q[sample], k)
        res["rep"] = (ri, rd, rep.info())
        return res

    outs = _ranks_in_threads(rank, world)
    for r, res in enumerate(outs):
        gi, gd, s = res[parallel.PN_EXCHANGE_ALLGATHER]
        assert np.array_equal(gi[sample].astype(np.uint64), oi.astype(np.uint64)) and np.array_equal(bits(gd[sample]), bits(od))
        assert s["nccl_bytes_sent"] > 0 and s["rows_out"] == nq
        gi, gd, s = res[parallel.PN_EXCHANGE_SLICE]
        lo, hi = parallel.query_slice(nq, r, world)
        assert gi.shape[0] == hi - lo and s["rows_out"] == hi - lo
        m = (sample >= lo) & (sample < hi)
        assert np.array_equal(gi[sample[m] - lo].astype(np.uint64), oi[m].astype(np.uint64)) and np.array_equal(bits(gd[sample[m] - lo]), bits(od[m]))
        ri, rd, inf = res["rep"]
        assert inf["device"] == r and inf["n_points"] == n
        assert np.array_equal(ri, oi.astype(np.uint64)) and np.array_equal(bits(rd), bits(od))
    [c.close() for c in comms]


@pytest.mark.parametrize("mode", [0, 1])
def test_multi_gpu_handle(pn, oracle, mode):
    """pn_multi_*: the whole box from one process through the C ABI (one GPU here means one rank; the 2+ GPU case runs when
    the box has them), host buffers in and out, both shard modes."""
    import torch
    from petal_neighbors_b200 import parallel, synth
    ng = torch.cuda.device_count()
    world = 1 << (min(ng, 8).bit_length() - 1)
    n, d, nq, k = 120000, 24, 90000, 10
    pts = synth.fast_gaussian_mixture(n, d, 9, n_centers=128, sigma=0.05)
    Q = synth.fast_gaussian_mixture(nq, d, 10, n_centers=128, sigma=0.05)
    m = parallel.MultiGpuBallTree(pts, list(range(world)), mode=mode)
    sample = np.arange(0, nq, nq // 300)[:300]
    oi, od = oracle.brute_knn(pts, Q[sample], k)
    # BY_SUBTREE: first the default exchange (merge kernels read the peers' lists over NVLink, no collective), then NCCL
    for exchange in ((None, parallel.PN_EXCHANGE_SLICE) if mode == 1 else (None,)):
        if exchange is not None:
            m.set_exchange(exchange)
        idx, dist = m.query_batch(Q, k)
        assert np.array_equal(idx[sample], oi.astype(np.uint64)) and np.array_equal(bits(dist[sample]), bits(od))
        st = m.stats()
        assert sum(s["rows_out"] for s in st) == nq
        if mode == 1 and world > 1:
            if exchange is None:
                assert all(s["nccl_calls"] == 0 and s["peer_mib"] > 0 for s in st)     # read from peer memory, no collective
            else:
                assert all(s["nccl_bytes_sent"] > 0 for s in st)
    m.close()
