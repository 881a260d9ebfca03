"""View sharding of one scene across the GPUs of a node.

The path shards by view: the encoder, the frame-attention blocks, every Linear / LayerNorm / MLP, the DPT and pose
heads, the decode and the post-processing are independent per view.  Only the global-attention blocks couple views;
there each rank keeps its own query rows and the K/V rows of all ranks are exchanged with one NCCL all-gather per
block (bf16, 2*D columns per token), issued asynchronously so that the attention over the LOCAL keys runs while the
remote keys are in flight; the second attention launch resumes the online softmax over the remote slots
(ma_attention_fwd_ex: kv segments + carried state).  The scale token lives on rank 0 (last row of its stream); its
final feature gives the metric scale, which is broadcast (one float).

Host-side logic only -- everything here also runs on CPU tensors with the gloo backend (tests/test_sharding_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_views(num_views: int, world: int) -> List[int]:
    """Contiguous view ranges in rank order, larger shards first (100 views on 8 ranks -> 13,13,13,13,12,12,12,12).
    Rank 0 always owns view 0 (the reference view)."""
    if num_views < world:
        raise ValueError(f"view sharding needs at least one view per rank (got {num_views} views for {world} ranks)")
    base, extra = divmod(num_views, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


@dataclass
class ViewShardPlan:
    """Row bookkeeping of the sharded info-sharing stream for one forward pass."""

    counts: Sequence[int]      # views per rank
    rank: int
    tokens_per_view: int       # N
    extra_tokens: int = 1      # the scale token, on rank 0

    @property
    def world(self) -> int:
        return len(self.counts)

    def rows(self, r: Optional[int] = None) -> int:
        """Token rows of rank r's residual stream (views * N, + the scale token on rank 0)."""
        r = self.rank if r is None else r
        return self.counts[r] * self.tokens_per_view + (self.extra_tokens if r == 0 else 0)

    @property
    def local_views(self) -> int:
        return self.counts[self.rank]

    @property
    def view_offset(self) -> int:
        return sum(self.counts[: self.rank])

    @property
    def total_views(self) -> int:
        return sum(self.counts)

    @property
    def total_rows(self) -> int:
        return sum(self.rows(r) for r in range(self.world))

    @property
    def slot_rows(self) -> int:
        """Rows of one rank's slot in the all-gather buffer (all slots equal: all_gather_into_tensor), 8-row aligned."""
        m = max(self.rows(r) for r in range(self.world))
        return (m + 7) // 8 * 8

    def local_segment(self) -> List[Tuple[int, int]]:
        return [(self.rank * self.slot_rows, self.rows())]

    def remote_segments(self) -> List[Tuple[int, int]]:
        """Other ranks' slots, starting with the next rank (spreads the first remote reads over different sources)."""
        order = [(self.rank + i) % self.world for i in range(1, self.world)]
        return [(r * self.slot_rows, self.rows(r)) for r in order if self.rows(r) > 0]


class ViewShardComm:
    """The collectives the sharded path needs, over a torch.distributed process group (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("view sharding needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_counts(self, local_views: int, device) -> List[int]:
        t = torch.zeros(self.world, dtype=torch.int64, device=device)
        mine = torch.tensor([local_views], dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(t, mine, group=self.group)
        return [int(x) for x in t.tolist()]

    def any_flags(self, flags: Sequence[bool], device) -> List[bool]:
        """Element-wise OR of per-rank booleans (all-reduce MAX): decisions that select a code path containing collectives
        -- "does ANY view of the scene provide this modality" -- must be taken on scene-wide facts, or ranks whose shard
        differs would issue different collectives and hang."""
        t = torch.tensor([1 if f else 0 for f in flags], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return [bool(x) for x in t.tolist()]

    def all_gather_slots(self, buf: torch.Tensor, slot_rows: int):
        """In-place all-gather of `buf` [world*slot_rows, C]: every rank has filled its own slot.  Returns a work handle;
        `.wait()` makes the CURRENT stream wait for the gather (no host block on NCCL)."""
        mine = buf[self.rank * slot_rows:(self.rank + 1) * slot_rows]
        return dist.all_gather_into_tensor(buf, mine, group=self.group, async_op=True)

    def all_gather_rows(self, rows: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
        """Concatenates every rank's [counts[r], C] rows in rank order (small per-view scalars: poses)."""
        m = max(counts)
        pad = rows.new_zeros(m, rows.shape[1])
        pad[: rows.shape[0]] = rows
        out = rows.new_empty(self.world * m, rows.shape[1])
        dist.all_gather_into_tensor(out, pad, group=self.group)
        return torch.cat([out[r * m:r * m + counts[r]] for r in range(self.world)], dim=0)

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
        return t
