"""One process per GPU (torchrun): volumes are sharded by rank exactly like the reference's
per-GPU subprocess fan-out (lib/utils/my_subprocess.py:56: np.array_split(range(N), NUM_GPUS));
the only exchange step gathers the per-volume detections and sums integer eval counts
(SURVEY.md section 8e).  NCCL on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_indices(n_items, world_size, rank):
    """Indices of the work units owned by `rank` (np.array_split semantics)."""
    return np.array_split(np.arange(n_items), world_size)[rank]


def gather_detections(local_dets, local_ids, group=None, device=None):
    """All-gather variable-length per-volume detections.

    local_dets: list of float32 arrays [n_i,7] (one per local volume); local_ids: their global
    volume ids.  Returns {volume_id: [n,7] array} on every rank.  Two collectives: one all_gather of
    the counts, one all_gather of a padded [max_rows, 8] block (col 0 = volume id)."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    rows = [np.hstack([np.full((d.shape[0], 1), i, dtype=np.float32), np.asarray(d, dtype=np.float32).reshape(-1, 7)])
            for d, i in zip(local_dets, local_ids)]
    block = np.concatenate(rows, axis=0) if rows else np.zeros((0, 8), dtype=np.float32)
    n_local = torch.tensor([block.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(ws)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    padded = torch.zeros((m, 8), dtype=torch.float32, device=device)
    if block.shape[0]:
        padded[:block.shape[0]] = torch.from_numpy(block).to(device)
    out = [torch.zeros((m, 8), dtype=torch.float32, device=device) for _ in range(ws)]
    dist.all_gather(out, padded, group=group)
    result = {}
    for r in range(ws):
        blk = out[r][:counts[r]].cpu().numpy()
        for vid in np.unique(blk[:, 0]).astype(np.int64) if blk.shape[0] else []:
            result[int(vid)] = blk[blk[:, 0] == vid][:, 1:].copy()
    return result


def allreduce_counts(counts, group=None, device=None):
    """SUM all-reduce of integer evaluation counters (tp, fp, n_pos, ...)."""
    import torch
    import torch.distributed as dist
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
