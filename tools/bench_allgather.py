"""How fast is the K/V all-gather of the sharded global attention, alone and under the local-key attention it is meant to hide
behind?  (torchrun --nproc-per-node N tools/bench_allgather.py [views_per_rank])

Prints one JSON line (rank 0): the NCCL all-gather of the [N * slot_rows, 2D] bf16 buffer
  alone      - nothing else on the GPU,
  overlapped - issued on NCCL's stream while ma_attention_fwd over the local keys runs on the compute stream (the real schedule
               of Engine._block_global_sharded), with the time the compute stream then waits for it,
and the duration of the local attention itself.
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))


def main():
    vpr = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mapanything_b200 import ops

    H, D = 12, 768
    rows = vpr * 1369 + (1 if rank == 0 else 0)
    slot = (vpr * 1369 + 1 + 7) // 8 * 8
    kv = torch.randn(world * slot, 2 * D, device=dev).bfloat16()
    q = torch.randn(rows, D, device=dev).bfloat16()
    so = torch.empty(1, rows, D, device=dev)
    sm = torch.full((1, rows, H), float("-inf"), device=dev)
    mine = kv[rank * slot:(rank + 1) * slot]

    def gather():
        return dist.all_gather_into_tensor(kv, mine, async_op=True)

    def local_attention():
        ops.attention(q, kv[:, :D], kv[:, D:], None, num_heads=H, num_seqs=1, q_len=rows, kv_len=rows,
                      kv_seq_stride=kv.shape[0], kv_segments=[(rank * slot, rows)], state=(so, sm), state_out=True)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for _ in range(3):
        gather().wait()
        local_attention()
    torch.cuda.synchronize()
    dist.barrier()
    res = {"world": world, "views_per_rank": vpr, "gather_MB_total": kv.numel() * 2 / 1e6}
    t = []
    for _ in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        a = ev()
        gather().wait()
        b = ev()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    res["allgather_alone_ms"] = sorted(t)[len(t) // 2]
    t = []
    for _ in range(5):
        torch.cuda.synchronize()
        a = ev()
        local_attention()
        b = ev()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    res["local_attention_ms"] = sorted(t)[len(t) // 2]
    tw, tt = [], []
    for _ in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        a = ev()
        w = gather()
        local_attention()
        b = ev()
        w.wait()
        c = ev()
        torch.cuda.synchronize()
        tw.append(b.elapsed_time(c))
        tt.append(a.elapsed_time(c))
    res["overlapped_exposed_wait_ms"] = sorted(tw)[len(tw) // 2]
    res["overlapped_total_ms"] = sorted(tt)[len(tt) // 2]
    out = torch.tensor([res["allgather_alone_ms"], res["local_attention_ms"], res["overlapped_exposed_wait_ms"],
                        res["overlapped_total_ms"]], device=dev)
    dist.all_reduce(out, op=dist.ReduceOp.MAX)
    if rank == 0:
        res.update(dict(zip(("allgather_alone_ms", "local_attention_ms", "overlapped_exposed_wait_ms", "overlapped_total_ms"),
                            [round(x, 4) for x in out.tolist()])))
        res["allgather_alone_busbw_GBs"] = res["gather_MB_total"] * (world - 1) / world / res["allgather_alone_ms"]
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
