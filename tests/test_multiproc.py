"""N>1 host logic on CPU: world_size-2 gloo process group (work partition + timing reductions)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tcvom_b200 import dp
    a, b = dp.shard_range(n_items, rank, world)
    rr = dp.shard_round_robin(n_items, rank, world)
    # every rank "processes" its windows: checksum = sum of the window ids it owns
    owned = torch.zeros(n_items)
    owned[a:b] += 1
    dist.all_reduce(owned)
    owned_rr = torch.zeros(n_items)
    owned_rr[rr] += 1
    dist.all_reduce(owned_rr)
    ms = 10.0 * (rank + 1)                      # rank 1 is slower
    thr = dp.job_throughput(b - a, ms)
    mx = dp.max_over_ranks(ms)
    dp.barrier()
    if rank == 0:
        out.put(dict(owned=owned.tolist(), owned_rr=owned_rr.tolist(), thr=thr, mx=mx))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8, 1])
def test_two_rank_partition_and_reductions(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["owned"] == [1.0] * n_items          # every window owned by exactly one rank
    assert res["owned_rr"] == [1.0] * n_items
    assert res["mx"] == 20.0                         # max over ranks, not rank 0's own time
    assert abs(res["thr"] - n_items / 0.020) < 1e-6  # all units / slowest rank's time


def test_single_process_paths():
    from tcvom_b200 import dp
    assert dp.shard_range(10, 0, 1) == (0, 10)
    assert dp.max_over_ranks(3.0) == 3.0
    assert dp.job_throughput(4, 2000.0) == 2.0
    # the reference's own partition rule (pred_test.py:125-131)
    assert [dp.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]


def test_install_hooks_reference_when_available():
    import sys
    ref = os.environ.get("TCVOM_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "models")):
        pytest.skip("reference checkout not present (GPU box)")
    sys.path.insert(0, ref)
    try:
        import tcvom_b200
        tcvom_b200.install()
        from models.model import EvalModel
        m = EvalModel("vmn_gca", agg_window=7, dilate_kernel=None)
        assert isinstance(m, tcvom_b200.EvalModel) and isinstance(m.NET, tcvom_b200.VMN)
        import models.VMN as V
        ref_net = V.get_VMN_models._reference("vmn_gca", 7)
        m.NET.load_state_dict(ref_net.state_dict(), strict=True)
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]


def test_install_native_tam_rebinds_the_reference_decoders():
    """install(native_tam=True): the FeatureAggregationModule name the reference decoders instantiate (VMN_DIM.py:4,99,
    VMN_Index.py:5,10, VMN_FBA.py:3,9, VMN_GCA.py:6,15) resolves to the native operator module; the networks the reference's
    own factory then builds carry it and still take the reference's state_dict."""
    import sys
    ref = os.environ.get("TCVOM_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "models")):
        pytest.skip("reference checkout not present (GPU box)")
    sys.path.insert(0, ref)
    try:
        import tcvom_b200
        import models.VMN as V
        from models.VMN.VMN_model import FeatureAggregationModule as RefTAM
        factory = getattr(V.get_VMN_models, "_reference", V.get_VMN_models)
        ref_net = factory("vmn_index", 7)
        assert type(ref_net.decoder.fam) is RefTAM
        tcvom_b200.install(native_tam=True)
        for name in ("VMN_model", "VMN_DIM", "VMN_Index", "VMN_FBA", "VMN_GCA"):
            mod = sys.modules["models.VMN." + name]
            assert mod.FeatureAggregationModule is tcvom_b200.FeatureAggregationModule, name
        assert tcvom_b200.FeatureAggregationModule._reference is RefTAM
        factory = V.get_VMN_models._reference
        for arch, chn in (("vmn_index", 32), ("vmn_dim", 256)):
            net = factory(arch, 7)
            assert type(net.decoder.fam) is tcvom_b200.FeatureAggregationModule
            assert net.decoder.fam.key_conv.weight.shape == (chn, chn, 3, 3)
            if arch == "vmn_index":
                net.load_state_dict(ref_net.state_dict(), strict=True)
    finally:
        sys.path.remove(ref)
        if hasattr(tcvom_b200.FeatureAggregationModule, "_reference"):
            del tcvom_b200.FeatureAggregationModule._reference
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
