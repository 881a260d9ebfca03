"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: view partitioning, the per-rank slot layout of the K/V
all-gather buffer, the collectives wrapper, and the two-launch "local keys first, remote keys later" online-softmax
protocol the sharded global attention uses (restated in torch here; the CUDA kernel is tested in test_attention_gpu.py
against the same semantics)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mapanything_b200.sharding import ViewShardComm, ViewShardPlan, partition_views


def test_partition_views():
    assert partition_views(100, 8) == [13, 13, 13, 13, 12, 12, 12, 12]
    assert partition_views(8, 8) == [1] * 8
    assert partition_views(24, 4) == [6, 6, 6, 6]
    assert sum(partition_views(1000, 8)) == 1000
    with pytest.raises(ValueError):
        partition_views(3, 4)


def test_plan_rows_and_segments():
    N = 1369
    counts = [13, 13, 12]
    plans = [ViewShardPlan(counts, r, N) for r in range(3)]
    assert [p.rows() for p in plans] == [13 * N + 1, 13 * N, 12 * N]
    assert all(p.slot_rows == plans[0].slot_rows and p.slot_rows % 8 == 0 and p.slot_rows >= 13 * N + 1 for p in plans)
    assert plans[0].total_rows == 38 * N + 1
    assert [p.view_offset for p in plans] == [0, 13, 26]
    for p in plans:
        segs = p.local_segment() + p.remote_segments()
        assert sorted(segs) == [(r * p.slot_rows, plans[r].rows()) for r in range(3)]
        assert p.remote_segments()[0][0] == ((p.rank + 1) % 3) * p.slot_rows  # next rank first
        assert sum(l for _, l in segs) == p.total_rows


def _attn_state(q, k, v, scale, state=None):
    """Torch restatement of ma_attention_fwd_ex's carried state: returns (normalised o, m') with
    m' = max + log2(sum) / (scale * log2 e) in raw-score units."""
    s = q @ k.t()
    sl2 = scale * math.log2(math.e)
    m = s.max(dim=1).values
    o_prev = None
    if state is not None:
        o_prev, m_prev = state
        m = torch.maximum(m, m_prev)
    p = torch.exp2((s - m[:, None]) * sl2)
    l = p.sum(1)
    o = p @ v
    if state is not None:
        a = torch.exp2((m_prev - m) * sl2)
        l = l + a
        o = o + o_prev * a[:, None]
    return o / l[:, None], m + torch.log2(l) / sl2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, counts, D, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ViewShardComm()
        got = comm.exchange_counts(counts[rank], torch.device("cpu"))
        assert got == list(counts)
        plan = ViewShardPlan(got, rank, N)
        # every rank can build the whole scene from the seed; it only "owns" its rows
        g = torch.Generator().manual_seed(0)
        total = plan.total_rows
        q_all = torch.randn(total, D, generator=g)
        kv_all = torch.randn(total, 2 * D, generator=g)
        # global row order: rank 0's views, the scale token (rank 0's last row), then the other ranks' views
        starts = [sum(plan.rows(r) for r in range(i)) for i in range(world)]
        lo, hi = starts[rank], starts[rank] + plan.rows()
        slot = plan.slot_rows
        kvbuf = torch.zeros(world * slot, 2 * D)
        kvbuf[rank * slot:rank * slot + plan.rows()] = kv_all[lo:hi]
        work = comm.all_gather_slots(kvbuf, slot)
        q = q_all[lo:hi]
        scale = D ** -0.5
        (r0, ln), = plan.local_segment()
        state = _attn_state(q, kvbuf[r0:r0 + ln, :D], kvbuf[r0:r0 + ln, D:], scale)  # overlaps the gather on a GPU
        work.wait()
        for r0, ln in plan.remote_segments():
            state = _attn_state(q, kvbuf[r0:r0 + ln, :D], kvbuf[r0:r0 + ln, D:], scale, state)
        ref = torch.softmax(q @ kv_all[:, :D].t() * scale, dim=1) @ kv_all[:, D:]
        err = (state[0] - ref).abs().max().item()
        s = torch.tensor([3.25 if rank == 0 else 0.0])
        comm.broadcast(s, src=0)
        # scene-wide modality flags: every rank must see the OR of all ranks' flags (selects paths with collectives)
        assert comm.any_flags([rank == 0, False, rank == 1], torch.device("cpu")) == [True, False, True]
        out_q.put((rank, err, float(s.item())))
    finally:
        dist.destroy_process_group()


def test_sharded_global_attention_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    N, counts, D = 37, (3, 2), 32
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, counts, D, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out_q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, s in res:
        assert err < 1e-5, (rank, err)
        assert s == 3.25
