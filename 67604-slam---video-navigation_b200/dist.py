"""Multi-GPU sharding: one process per GPU, torch.distributed for the plumbing.

The front-end shards without any data-path exchange:
  * sequence (configs 2/3): contiguous blocks of frame pairs per rank, balanced by descriptor-pair
    work; a rank also runs the stereo stage of its right halo frame;
  * loop closure (config 4): the lower-triangular (keyframe, earlier keyframe) candidate list is
    cut into blocks balanced by Nq*Nt;
  * dense sweep (config 5): the TRAIN set is cut into slices, queries are replicated, every rank
    emits per-query top-2 keys with GLOBAL train indices.
The only collective is one all-gather of the per-shard result tables (NCCL over NVLink on GPUs,
gloo in the CPU tests); the top-2 tables are then min-merged (exact, keys are totally ordered).
The partitioning / merge logic below is pure host code and backend-agnostic.
"""
from __future__ import annotations

import os

import numpy as np


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        import datetime
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        # a mismatched collective must fail in minutes, not hold N GPUs for the default 10 minutes
        timeout = datetime.timedelta(seconds=int(os.environ.get("SLAMFE_PG_TIMEOUT_S", "120")))
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank), timeout=timeout)
        else:
            dist.init_process_group(backend, timeout=timeout)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world, local_rank


def bind_to_gpu_numa(gpu_index):
    """Restrict this process to the CPUs NVML reports as local to its GPU (best effort, Linux only), so
    that pinned staging buffers are first-touched on the GPU's NUMA node.  Returns the previous affinity
    set (pass it to os.sched_setaffinity(0, ...) to undo) or None when nothing was changed.  Opt-in:
    bench.py calls it when SLAMFE_BIND_NUMA=1 (to be measured at 8 GPUs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        old = os.sched_getaffinity(0)
        new = cpus & old
        if not new or new == old:
            return None
        os.sched_setaffinity(0, new)
        return old
    except Exception:
        return None


def balanced_ranges(work, world):
    """Cut len(work) consecutive units into `world` contiguous ranges of near-equal total work.
    Returns an int64 array of world+1 boundaries."""
    work = np.asarray(work, dtype=np.float64)
    n = len(work)
    bounds = np.zeros(world + 1, dtype=np.int64)
    if n == 0:
        return bounds
    csum = np.concatenate([[0.0], np.cumsum(work)])
    total = csum[-1]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        if b > 0 and b <= n and target - csum[b - 1] < csum[b] - target:
            b -= 1  # the boundary whose prefix sum is nearest to the target
        # keep boundaries monotone and leave at least zero units per rank
        bounds[r] = min(max(b, bounds[r - 1]), n)
    bounds[world] = n
    return bounds


def frame_pair_shards(n_l, n_r, world):
    """Shard frames by stereo work Nl*Nr.  Returns boundaries b: rank r owns frames
    [b[r], b[r+1]) and the frame pairs (f, f+1) for f in that range (f+1 < n_frames)."""
    n_l, n_r = np.asarray(n_l, dtype=np.int64), np.asarray(n_r, dtype=np.int64)
    return balanced_ranges(n_l * n_r, world)


def candidate_pairs(n_keyframes):
    """All (keyframe i, earlier keyframe j < i) pairs, row-major: the brute-force superset of
    loop_closure.py:315-348's gated candidates (BASELINE config 4)."""
    i, j = np.tril_indices(n_keyframes, k=-1)
    return np.stack([i, j], axis=1).astype(np.int32)


def candidate_blocks(pairs, sizes, world):
    """Cut the candidate pair list into `world` contiguous blocks balanced by Nq*Nt."""
    sizes = np.asarray(sizes, dtype=np.int64)
    return balanced_ranges(sizes[pairs[:, 0]] * sizes[pairs[:, 1]], world)


def train_slices(n_train, world, align=16):
    """Boundaries of the train-set slices for the dense sweep (multiples of `align` rows so every
    slice starts 16-byte aligned in the 61-byte-row layout)."""
    per = -(-n_train // world)
    per = -(-per // align) * align
    b = np.minimum(np.arange(world + 1, dtype=np.int64) * per, n_train)
    return b


def merge_top2_host(shard_keys):
    """Host restatement of slamfe_merge_top2 on uint32 keys (n_shards, nq, 2) -> (nq, 2)."""
    k = np.asarray(shard_keys).view(np.uint32)
    flat = np.concatenate([k[s] for s in range(k.shape[0])], axis=1)
    flat.sort(axis=1)
    return np.ascontiguousarray(flat[:, :2])


def _all_gather_stacked(t):
    """all_gather_into_tensor with a (world, ...) result; the flat (world*n, ...) output form is
    the one both NCCL and gloo accept."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = t.contiguous()
    if t.dim() == 0:
        t = t.reshape(1)
    flat = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(flat, t)
    return flat.view((world, t.shape[0]) + tuple(t.shape[1:]))


def all_gather_padded(t, lengths=None):
    """All-gather a per-rank tensor whose leading dimension differs between ranks.

    Returns (gathered (world, max_len, ...), lengths (world,)) — fixed-stride tables, one
    collective for the payload (plus one tiny one for the lengths when they are not given)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t.unsqueeze(0), np.array([t.shape[0]], dtype=np.int64)
    world = dist.get_world_size()
    if lengths is None:
        ln = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        lengths = _all_gather_stacked(ln).reshape(-1).cpu().numpy()
    max_len = int(max(lengths))
    pad = torch.zeros((max_len,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    return _all_gather_stacked(pad), np.asarray(lengths, dtype=np.int64)


def gather_and_merge_top2(row_keys):
    """All-gather every rank's (nq, 2) key table (keys carry GLOBAL train indices) and min-merge
    them: exact global top-2 on every rank.  CUDA tensors merge with slamfe_merge_top2 over NCCL,
    CPU tensors (gloo, used by the host-logic tests) with the host restatement."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return row_keys
    gathered = _all_gather_stacked(row_keys)
    if row_keys.is_cuda:
        from . import ops
        return ops.merge_top2(gathered)
    return torch.from_numpy(merge_top2_host(gathered.numpy()).view(np.int32))


def sharded_knn(q_dev, t_shard_dev, t_index_base, desc_bytes=None, best_only=False):
    """Dense sweep (config 5): this rank's train slice vs all queries, all-gather, exact merge.
    Returns the global (nq, 2) key table on every rank."""
    from . import ops
    row_keys, _ = ops.hamming_top2(q_dev, t_shard_dev, desc_bytes=desc_bytes, t_index_base=int(t_index_base),
                                   best_only=best_only)
    return gather_and_merge_top2(row_keys)


def gather_pair_tables(tables, n_units):
    """All-gather per-shard result tables of the frame-pair / candidate-block shardings.
    `tables` maps name -> tensor whose leading dimension is this rank's unit count (frames, candidate
    pairs) or any other per-rank length (rows).  Returns name -> (world, max_len, ...) plus the per-rank
    unit counts.  The true leading length of EVERY table is exchanged first (one small collective), so
    ranks whose rows-per-unit differ still agree on the padded size."""
    import torch
    import torch.distributed as dist
    first = next(iter(tables.values()))
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: v.unsqueeze(0) for k, v in tables.items()}, np.array([n_units], dtype=np.int64)
    names = sorted(tables)
    ln = torch.tensor([n_units] + [tables[k].shape[0] for k in names], dtype=torch.int64, device=first.device)
    all_len = _all_gather_stacked(ln).cpu().numpy()          # (world, 1 + n_tables)
    out = {}
    for i, k in enumerate(names):
        out[k], _ = all_gather_padded(tables[k], lengths=all_len[:, 1 + i])
    return out, all_len[:, 0].astype(np.int64)
