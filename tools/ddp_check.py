"""2-GPU check of the native training step under the reference's own distributed recipe (train_ddp.py:270-280):
SyncBatchNorm.convert_sync_batchnorm -> DistributedDataParallel(find_unused_parameters=True), NCCL.

Rank r trains on sample r of tests/golden/train_step_s5.npz.  Expected values: forward quantities (alpha, BatchNorm
running statistics, spectral-norm u/v) are those of the reference's single-process B=2 step (SyncBN makes them
identical); gradients are the oracle's DDP emulation (global BN statistics, per-rank losses, mean over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tcvom_b200  # noqa: E402
from helpers import fixture_sd, golden, key_table  # noqa: E402
from oracle import vmn_gca_oracle as O  # noqa: E402

LOSS_WEIGHTS = (1.0, 1.0, 1.0, 0.5, 0.25)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", init_method="env://")
    dev = torch.device("cuda", local)
    g = golden("train_step_s5.npz")
    assert world == g["a"].shape[0], "one golden sample per rank"
    model = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3)
    model.NET.load_state_dict(fixture_sd(), strict=True)
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).to(dev)
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local,
                                                      find_unused_parameters=True)
    model.train()
    a, fg, bg = (torch.from_numpy(g[k][rank:rank + 1]).float().to(dev) for k in ("a", "fg", "bg"))
    out = model(a, fg, bg)
    loss = sum(w * o.mean() for w, o in zip(LOSS_WEIGHTS, out[:5]))
    model.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    net = model.module.NET
    res = {"rank": rank}
    res["alpha_err"] = float((out[7].detach().cpu() - torch.from_numpy(g["alphas"][rank:rank + 1])).abs().max())
    sd = net.state_dict()
    st_err = 0.0
    for k in g.files:
        if k.startswith("st:"):
            ref = torch.from_numpy(g[k]).float()
            st_err = max(st_err, float((sd[k[3:]].detach().cpu().float() - ref).abs().max()) / max(1.0, float(ref.abs().max())))
    res["state_err"] = st_err
    # every rank must hold the same (averaged) gradients
    named = dict(net.named_parameters())
    trainable = key_table()["trainable"]
    flat = torch.cat([named[n].grad.flatten() for n in trainable])
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    res["rank_spread"] = float(max((o - flat).abs().max() for o in other))
    if rank == 0:
        osd = {k: v.clone() for k, v in fixture_sd().items()}
        for n in trainable:
            osd[n].requires_grad_(True)
        full = [torch.from_numpy(g[k]).float() for k in ("a", "fg", "bg")]
        oo = O.full_vmd_forward(osd, *full, [3] * world, train=True, rank_rows=[slice(r, r + 1) for r in range(world)])
        total = sum(sum(w * l for w, l in zip(LOSS_WEIGHTS, [oo[i][r] for i in range(5)])) for r in range(world)) / world
        total.backward()
        errs = []
        for n in trainable:
            ref = osd[n].grad if osd[n].grad is not None else torch.zeros_like(osd[n])
            got = named[n].grad.detach().cpu()
            errs.append((float((got.double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-12), n))
        errs.sort(reverse=True)
        res["grad_worst"], res["grad_worst_name"] = errs[0]
        res["grad_median"] = errs[len(errs) // 2][0]
        res["loss_native"] = [float(o) for o in out[:5]]
        res["loss_oracle_rank0"] = [float(oo[i][0]) for i in range(5)]
        res["sync_bn_modules"] = sum(isinstance(m, torch.nn.SyncBatchNorm) for m in net.modules())
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        ok = all(r["alpha_err"] < 1e-3 and r["state_err"] < 1e-3 and r["rank_spread"] == 0.0 for r in gathered) and \
            res["grad_worst"] < 1e-1 and res["grad_median"] < 4e-2
        print(json.dumps(dict(ok=ok, ranks=gathered)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
