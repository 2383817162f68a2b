"""2-GPU check of the SyncBatchNorm statistic exchange over NVLink peer memory (tcv_peer_allreduce_f64, tcvom_b200/peer.py)
against NCCL:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/peer_check.py
Prints one JSON line on rank 0: ok, max difference to the NCCL sum, bit-identity across ranks, microseconds per call of
both transports (CUDA events, 200 back-to-back calls of a 2 x 512 x 2 block -- the SyncBN message of a 512-channel layer)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcvom_b200.peer import make_peer_reducer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    red = make_peer_reducer(None, dev)
    res = dict(rank=rank, world=world, available=red is not None)
    if red is not None:
        st = torch.cuda.current_stream(dev)
        worst, same = 0.0, True
        for i, n in enumerate((1, 7, 64, 1024, 5120, 8192, 3, 2048)):        # odd sizes, both slots, the maximum
            g = torch.Generator().manual_seed(100 * i + rank)
            x = torch.randn(n, generator=g, dtype=torch.float64).to(dev) * (10.0 ** (i % 4))
            ref = x.clone()
            dist.all_reduce(ref)
            red.allreduce_(x, st.cuda_stream)
            torch.cuda.synchronize(dev)
            worst = max(worst, float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-300)))
            gathered = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(gathered, x)
            same = same and all(torch.equal(gathered[0], t) for t in gathered)
        res.update(max_rel_diff_to_nccl=worst, identical_across_ranks=same)
        x = torch.randn(2 * 512 * 2, dtype=torch.float64, device=dev)
        for name, fn in (("peer_us", lambda: red.allreduce_(x, st.cuda_stream)), ("nccl_us", lambda: dist.all_reduce(x))):
            for _ in range(20):
                fn()
            torch.cuda.synchronize(dev)
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
                x.mul_(0.5)                     # a dependent kernel between the calls, like bn_finalize in the real step
            e1.record()
            torch.cuda.synchronize(dev)
            res[name] = 1e3 * e0.elapsed_time(e1) / 200
        res["ok"] = bool(worst < 1e-14 and same)
    else:
        res["ok"] = False
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
