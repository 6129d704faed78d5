"""Multi-process check of the MCU-row sharded encoder (run under torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py [W H family]
Every rank generates its own rows of one synthetic image and runs jpezy_b200.shard.ShardedEncoder (NCCL all-gathers, P2P
stores into rank 0's buffer); rank 0 compares the stitched stream with the single-GPU encode of the whole image."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpezy_b200 as J  # noqa: E402
from jpezy_b200 import shard  # noqa: E402


def main():
    W, H, fam = (int(x) for x in (sys.argv[1:4] + ["4096", "2176", "1"][len(sys.argv) - 1:]))
    rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = J.Context(lr)
    row0, nrows = shard.partition_mcu_rows((H + 15) // 16, world)[rank]
    y0, ny = shard.pixel_rows(H, row0, nrows)
    planes = torch.empty((3, ny, W), dtype=torch.uint8, device="cuda")
    # a real (non-default) stream: the C ABI reads a NULL stream as "the context's own stream", which the NCCL collectives
    # of torch.distributed would not be ordered with
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    assert st != 0
    ctx.synth_rows_dev(planes[0], planes[1], planes[2], W, y0, ny, frame=0, family=fam, stream=st)
    enc = shard.ShardedEncoder(ctx, shard.DistGroup(dist, torch.device("cuda", lr)), max(W * H, 1 << 20))
    ok = True
    for it in range(3):
        enc.encode(planes[0], planes[1], planes[2], W, H, row0, nrows, y0, False, stream=st)
        seg, bits = enc.result()
        if rank == 0:
            whole = torch.empty((3, H, W), dtype=torch.uint8, device="cuda")
            ctx.synth_dev(whole[0], whole[1], whole[2], W, H, 1, 0, fam, stream=st)
            out = torch.zeros(max(W * H, 1 << 20), dtype=torch.uint8, device="cuda")
            nb = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.encode_batch_dev(whole[0], whole[1], whole[2], W, H, 1, False, out, out.numel(), nb, None, stream=st)
            torch.cuda.synchronize()
            want = out[: int(nb.item())].cpu().numpy().tobytes()
            ok = ok and seg == want
            print("iteration %d: %d ranks, %dx%d family %d: stitched %d bytes, single-GPU %d bytes, identical=%s, bits/rank=%s" %
                  (it, world, W, H, fam, len(seg), len(want), seg == want, bits), flush=True)
    enc.close()
    dist.destroy_process_group()
    ctx.close()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
