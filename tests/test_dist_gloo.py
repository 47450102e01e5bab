"""Multi-rank host logic on CPU: world_size 2, gloo backend (no GPU needed).

The product has no CPU compute path, so each rank's LOCAL result tables are produced by the oracle
here (test infrastructure); what is under test is the sharding arithmetic, the single all-gather
and the exact min-merge of packed keys with global indices (slamfe.dist)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(idx, dist_):
    k = (dist_.astype(np.uint32) << 22) | idx.astype(np.uint32)
    k[idx < 0] = 0xFFFFFFFF
    return k


def _worker(rank, world, port, q_out):
    try:
        _worker_body(rank, world, port, q_out)
    except Exception as exc:  # surface the failure instead of letting the parent time out
        import traceback
        q_out.put((rank, {"exception: " + repr(exc) + traceback.format_exc(): False}))
        raise


def _worker_body(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as tdist
    import slamfe  # noqa: F401
    from slamfe import dist as sdist, synth
    from oracle import ref_oracle as ora
    r, w, _ = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    res = {}
    # ---- dense sweep (config 5): train slices, global index bases, all-gather, exact merge ----
    rng = np.random.default_rng(5)  # same inputs on every rank
    q = synth.descriptors(rng, 333)
    t = synth.paired_descriptors(rng, q, n_out=777, dup_frac=0.1)[0]
    b = sdist.train_slices(len(t), world)
    idx2, d2 = ora.knn2(q, t[b[rank]:b[rank + 1]])
    local = np.stack([_pack(np.where(idx2[:, c] >= 0, idx2[:, c] + b[rank], -1), d2[:, c]) for c in (0, 1)], axis=1)
    merged = sdist.gather_and_merge_top2(torch.from_numpy(local.view(np.int32)))
    full_i, full_d = ora.knn2(q, t)
    want = np.stack([_pack(full_i[:, c], full_d[:, c]) for c in (0, 1)], axis=1)
    res["dense"] = bool(np.array_equal(merged.numpy().view(np.uint32), want))
    # ---- sequence (config 2): contiguous frame blocks + halo, gather of per-frame tables ----
    n_l = rng.integers(50, 120, 9); n_r = rng.integers(50, 120, 9)
    fb = sdist.frame_pair_shards(n_l, n_r, world)
    mine = np.arange(fb[rank], fb[rank + 1])
    counts = torch.tensor([int(n_l[f] * n_r[f]) for f in mine], dtype=torch.int64)
    gathered, lengths = sdist.all_gather_padded(counts)
    flat = np.concatenate([gathered[k, :lengths[k]].numpy() for k in range(world)])
    res["sequence"] = bool(np.array_equal(flat, n_l * n_r) and lengths.sum() == 9 and fb[0] == 0 and fb[-1] == 9)
    # ---- loop closure (config 4): candidate blocks balanced by Nq*Nt, gather of pair tables ----
    sizes = rng.integers(10, 60, 7)
    pairs = sdist.candidate_pairs(7)
    cb = sdist.candidate_blocks(pairs, sizes, world)
    pool = [synth.descriptors(rng, int(n)) for n in sizes]
    best = []
    for i, j in pairs[cb[rank]:cb[rank + 1]]:
        ti, td = ora.match(pool[i], pool[j])
        best.append(int(td.min()))
    tabs, lens = sdist.gather_pair_tables({"best": torch.tensor(best, dtype=torch.int32)}, len(best))
    allbest = np.concatenate([tabs["best"][k, :lens[k]].numpy() for k in range(world)])
    ref = np.array([int(ora.match(pool[i], pool[j])[1].min()) for i, j in pairs])
    res["loop"] = bool(np.array_equal(allbest, ref) and lens.sum() == len(pairs))
    # row tables whose rows-per-unit differ between ranks: the true lengths are exchanged
    rows_mine = [np.full(int(sizes[i]), 100 * i + j, np.int32) for i, j in pairs[cb[rank]:cb[rank + 1]]]
    tab = torch.from_numpy(np.concatenate(rows_mine) if rows_mine else np.zeros(0, np.int32))
    tabs2, lens2 = sdist.gather_pair_tables({"rows": tab, "best": torch.tensor(best, dtype=torch.int32)}, len(best))
    want_rows = [np.concatenate([np.full(int(sizes[i]), 100 * i + j, np.int32) for i, j in pairs[cb[k]:cb[k + 1]]])
                 for k in range(world)]
    res["ragged_rows"] = bool(all(np.array_equal(tabs2["rows"][k, :len(want_rows[k])].numpy(), want_rows[k])
                                  for k in range(world)) and np.array_equal(lens2, lens)
                              and tabs2["rows"].shape[1] == max(len(w_) for w_ in want_rows))
    work = (sizes[pairs[:, 0]] * sizes[pairs[:, 1]]).astype(np.float64)
    share = np.array([work[cb[k]:cb[k + 1]].sum() for k in range(world)]) / work.sum()
    res["balanced"] = bool(abs(share[0] - 0.5) < 0.15)
    tdist.barrier()
    tdist.destroy_process_group()
    q_out.put((rank, res))


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert all(res.values()), (rank, res)
