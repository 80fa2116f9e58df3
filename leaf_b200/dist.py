"""Multi-GPU sharding of the attack. One process per GPU, torch.distributed for the plumbing.

The reference has no working multi-GPU path for this step (its DDP wrapper does not forward encode_text),
so these modes are new design; what they must preserve is the single-process result:

* sample-sharded   rank r attacks samples [B*r/G, B*(r+1)/G) with all their candidates. The per-sample argmax is
                   local, so there is NO collective on the data path; one all-gather of the B winners' (z*, c*) pairs
                   and features at the end of a round makes every rank hold the global result.
* candidate-sharded rank r scores candidates [n*r/G, n*(r+1)/G) of EVERY sample; each phase ends with an all-gather
                   of (best loss f32, global candidate index) per sample - 8*B bytes per rank - and a local reduction
                   with torch.argmax's first-index tie-break on the GLOBAL candidate index.

Draws are made for the whole batch on every rank from identically seeded numpy RNGs, so the union of the shards
equals the single-GPU run exactly.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(total: int, rank: int, world_size: int):
    """Contiguous, balanced [lo, hi) of `total` items for `rank`."""
    return total * rank // world_size, total * (rank + 1) // world_size


def all_gather_cat(t: torch.Tensor, sizes, group=None) -> torch.Tensor:
    """Concatenate per-rank tensors with (possibly) different leading sizes `sizes[r]` along dim 0."""
    _, G = world(group)
    if G == 1:
        return t
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(G)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: sizes[r]] for r, o in enumerate(out)], dim=0)


def cross_shard_argmax(best_val: torch.Tensor, best_idx: torch.Tensor, group=None):
    """Per-sample argmax across candidate shards. best_val [B] fp32 = the local maximum score, best_idx [B] = its
    GLOBAL candidate index. Returns the global (value, index): the maximal value, ties broken by the smallest global
    index - exactly what torch.argmax over the unsharded [B, n] scores returns (utils_attacks.py:348,386)."""
    _, G = world(group)
    if G == 1:
        return best_val, best_idx
    vals = [torch.empty_like(best_val) for _ in range(G)]
    idxs = [torch.empty_like(best_idx) for _ in range(G)]
    dist.all_gather(vals, best_val.contiguous(), group=group)
    dist.all_gather(idxs, best_idx.contiguous(), group=group)
    v = torch.stack(vals, dim=1)                      # [B, G]
    i = torch.stack(idxs, dim=1)
    vmax = v.max(dim=1, keepdim=True).values
    cand = torch.where(v == vmax, i, torch.full_like(i, torch.iinfo(i.dtype).max))
    gi = cand.min(dim=1).values
    return vmax.squeeze(1), gi


def broadcast_rows(t: torch.Tensor, owner_of_row: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank holds t [B, ...] where only the rows it owns are meaningful (owner_of_row[b] = rank); returns the
    tensor with every row taken from its owner (sum of masked copies)."""
    rank, G = world(group)
    if G == 1:
        return t
    mask = (owner_of_row == rank).view((-1,) + (1,) * (t.dim() - 1)).to(t.dtype)
    out = t * mask
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
