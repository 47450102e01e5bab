"""Multi-rank parity on real NCCL (skipped on a box with fewer than 2 GPUs): the product's CUDA path per
rank, one all-gather over NVLink, slamfe_merge_top2 / candidate-block tables — against the oracle.
The host logic of the same calls is covered on CPU by tests/test_dist_gloo.py; the driver's scaling
run additionally checks the same parity inside bench.py's sub-records at N = 2, 4, 8."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q_out):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import torch
        import torch.distributed as tdist
        import slamfe  # noqa: F401
        from slamfe import dist as sdist, loop, synth
        from oracle import ref_oracle as ora
        r, w, lr = sdist.init_from_env(backend="nccl")
        dev = torch.device("cuda", lr)
        res = {}
        rng = np.random.default_rng(6)                     # same inputs on every rank
        # ---- dense sweep: this rank's train slice on its GPU, all-gather, slamfe_merge_top2 ----
        q = synth.descriptors(rng, 1500)
        t = synth.paired_descriptors(rng, q, n_out=2600, dup_frac=0.1)[0]
        b = sdist.train_slices(len(t), world)
        merged = sdist.sharded_knn(torch.from_numpy(q).to(dev), torch.from_numpy(t[b[rank]:b[rank + 1]]).to(dev), b[rank])
        oi, od = ora.knn2(q, t)
        k = merged.cpu().numpy().view(np.uint32)
        res["dense_top2_bit_exact"] = bool(np.array_equal(k & 0x3FFFFF, oi.astype(np.uint32)) and
                                           np.array_equal(k >> 22, od.astype(np.uint32)))
        # ---- loop closure: candidate blocks per rank, gathered best-match tables ----
        K, n = 9, 300
        pool = np.concatenate([synth.descriptors(rng, n) for _ in range(K)])
        links = synth.links(rng, K * n)
        pairs = sdist.candidate_pairs(K)
        cb = sdist.candidate_blocks(pairs, np.full(K, n), world)
        mine = pairs[cb[rank]:cb[rank + 1]]
        ver = loop.CandidateVerifier(block_pairs=16)
        table = torch.empty((len(mine) * n,), dtype=torch.int32, device=dev)
        ver.verify(torch.from_numpy(pool).to(dev), torch.from_numpy(links).to(dev), np.arange(K) * n, np.full(K, n), mine,
                   n_iter=16, seed=1, pair_base=int(cb[rank]), key_table=table, sync=False)
        gathered, lens = sdist.all_gather_padded(table, lengths=(cb[1:] - cb[:-1]) * n)
        g = gathered.cpu().numpy().view(np.uint32)
        ok = True
        for rr in range(world):
            for kk, (i, j) in enumerate(pairs[cb[rr]:cb[rr + 1]]):
                ti, td = ora.match(pool[i * n:(i + 1) * n], pool[j * n:(j + 1) * n])
                row = g[rr, kk * n:(kk + 1) * n]
                ok &= bool(np.array_equal(row & 0x3FFFFF, ti.astype(np.uint32)) and np.array_equal(row >> 22, td.astype(np.uint32)))
        res["loop_tables_bit_exact"] = ok
        tdist.barrier()
        tdist.destroy_process_group()
        q_out.put((rank, res))
    except Exception as exc:
        import traceback
        q_out.put((rank, {"exception: " + repr(exc) + traceback.format_exc(): False}))
        raise


def test_world_size_2_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert all(res.values()), (rank, res)
