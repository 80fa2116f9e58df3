"""Multi-GPU sharding of the attack. One process per GPU, torch.distributed for the plumbing.

The reference has no working multi-GPU path for this step (its DDP wrapper does not forward encode_text),
so these modes are new design; what they must preserve is the single-process result:

* sample-sharded   rank r attacks samples [B*r/G, B*(r+1)/G) with all their candidates. The per-sample argmax is
                   local, so there is NO collective on the data path; one all-gather of the B winners' (z*, c*) pairs
                   and features at the end of a round makes every rank hold the global result.
* candidate-sharded rank r scores candidates [n*r/G, n*(r+1)/G) of EVERY sample; each phase ends with ONE all-gather
                   of the packed (best loss f32, global candidate index i32) per sample - 8*B bytes per rank - and a
                   local reduction with torch.argmax's first-index tie-break on the GLOBAL candidate index. A rank
                   with no candidates (n < world size) reports NO_CANDIDATE and never wins.

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


NO_CANDIDATE = 0x7FFFFFFF          # global index an empty shard reports (never wins against a real candidate)


def pack_score_index(val: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(loss f32, global candidate index i32)[B] as ONE int64 per sample: the 8*B bytes a rank contributes per phase
    (SURVEY.md 8e). High word = the float's bits, low word = the index."""
    bits = val.to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return (bits << 32) | (idx.to(torch.int64) & 0xFFFFFFFF)


def unpack_score_index(packed: torch.Tensor):
    val = (packed >> 32).to(torch.int32).view(torch.float32)
    idx = (packed & 0xFFFFFFFF).to(torch.int64)
    return val, idx


def cross_shard_argmax(best_val: torch.Tensor, best_idx: torch.Tensor, group=None):
    """Per-sample argmax across candidate shards. best_val [B] fp32 = the local maximum score, best_idx [B] = its
    GLOBAL candidate index (NO_CANDIDATE for a rank that scored nothing). ONE all-gather of the packed pairs per phase;
    returns the global (value, index): the maximal value, ties broken by the smallest global index - exactly what
    torch.argmax over the unsharded [B, n] scores returns (utils_attacks.py:348,386). A NaN score never wins over a
    number (as in leaf_score's own argmax); if every shard reports NaN the smallest index is taken."""
    _, G = world(group)
    if G == 1:
        return best_val, best_idx
    mine = pack_score_index(best_val, best_idx)
    mine = mine.reshape(-1)
    out = torch.empty((G * mine.numel(),), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    v, i = unpack_score_index(out.view(G, -1).transpose(0, 1).contiguous())       # [B, G]
    none = torch.full_like(i, NO_CANDIDATE)
    real = ~torch.isnan(v) & (i != NO_CANDIDATE)
    v = torch.where(real, v, torch.full_like(v, float("-inf")))
    vmax = v.max(dim=1, keepdim=True).values
    gi = torch.where(real & (v == vmax), i, none).min(dim=1).values
    gi = torch.where(gi == NO_CANDIDATE, i.min(dim=1).values, gi)                 # every shard NaN: the smallest index
    return vmax.squeeze(1), gi


def broadcast_rows(t: torch.Tensor, owner_of_row: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank holds t [B, ...] where only the rows it owns are meaningful (owner_of_row[b] = rank); returns the
    tensor with every row taken from its owner (sum of copies in which the other rows are zero - selected, not
    multiplied, so a NaN/Inf in a row a rank does not own cannot leak)."""
    rank, G = world(group)
    if G == 1:
        return t
    mask = (owner_of_row == rank).view((-1,) + (1,) * (t.dim() - 1))
    out = torch.where(mask, t, torch.zeros_like(t))
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
