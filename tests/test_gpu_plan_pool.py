"""GPU: recorded plans allocate from a private memory pool (GcaVmnEngine.recording): dead intermediates are reused inside
the plan, the pool keeps the addresses reserved for the replays.  Checks that the pooled plan computes exactly what the
keep-everything plan computes, survives allocator churn between replays, and needs far less memory."""
import gc

import pytest
import torch

from helpers import fixture_sd

pytestmark = pytest.mark.gpu


def _model(pool: bool):
    import tcvom_b200
    m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=2)
    m.NET.load_state_dict(fixture_sd(), strict=True)
    m = m.cuda().eval()
    m.NET.engine().plan_pool = pool
    return m


def test_pooled_plan_equals_unpooled_and_survives_allocator_churn():
    from tcvom_b200 import synthetic
    H, W = 256, 384
    imgs, tris = synthetic.make_window(H, W, seed=11, frames=5)
    ti, tt = torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()
    outs, peaks = {}, {}
    for pool in (False, True):
        gc.collect(); torch.cuda.empty_cache(); torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        torch.cuda.reset_peak_memory_stats()
        m = _model(pool)
        a0 = m(ti, tt).clone()
        plan = m._plan(1, 5, H, W, ti.device, True)
        assert (plan.pool is not None) == pool
        peaks[pool] = torch.cuda.max_memory_allocated() - base
        # allocator churn between replays: release every cached block, then fill fresh allocations with garbage
        torch.cuda.empty_cache()
        junk = [torch.full((64 << 20,), float("nan"), device="cuda") for _ in range(8)]     # 2 GB
        torch.cuda.synchronize()
        a1 = m(ti, tt).clone()
        del junk
        torch.cuda.empty_cache()
        a2 = m(ti, tt).clone()
        assert torch.equal(a0, a1) and torch.equal(a0, a2)
        assert torch.isfinite(a0).all() and float(a0[:, 1:4].max()) > 0.1
        outs[pool] = a0
        del m, plan
    assert torch.equal(outs[False], outs[True])
    assert peaks[True] < 0.6 * peaks[False], peaks


def test_pooled_plan_memory_is_released_with_the_plan():
    from tcvom_b200 import synthetic
    H, W = 256, 384
    imgs, tris = synthetic.make_window(H, W, seed=12)
    ti, tt = torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()
    gc.collect(); torch.cuda.empty_cache(); torch.cuda.synchronize()
    m = _model(True)
    m(ti, tt)
    torch.cuda.synchronize()
    held = torch.cuda.memory_reserved()
    from tcvom_b200.engine import release_idle_pools
    m.NET.engine().plans.clear()
    gc.collect()
    assert release_idle_pools() >= 1               # the dead plan's pool was parked, not destroyed by the finaliser
    torch.cuda.empty_cache()
    assert torch.cuda.memory_reserved() < held


def test_many_models_recorded_back_to_back_reuse_pools():
    """Plans of dead models are finalised by the garbage collector at arbitrary moments -- also while another plan is
    being recorded (destroying a MemPool there aborts the process)."""
    from tcvom_b200 import synthetic
    from tcvom_b200.engine import _IDLE_POOLS
    imgs, tris = synthetic.make_window(64, 96, seed=3)
    ti, tt = torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()
    first = None
    for i in range(8):
        m = _model(True)
        holder = [m]
        holder.append(holder)                      # a reference cycle: only the cyclic collector frees this model
        a = m(ti, tt).clone()
        first = a if first is None else first
        assert torch.equal(a, first)
        del m, holder
    gc.collect()
    assert sum(len(v) for v in _IDLE_POOLS.values()) >= 1
