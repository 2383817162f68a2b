"""GPU: the native training path (train-mode forward + backward kernels behind the C ABI).

Operator level: every tape operator of tcvom_b200.train_engine against torch autograd / the CPU oracle on
seeded inputs (tools/train_check.py holds the checks).  Step level: one FullModel_VMD training step against one
step of the UNMODIFIED reference (tests/golden/train_step_s5.npz): losses, alphas, all 228 gradients, the
spectral-norm u/v and BatchNorm running statistics the forward mutates."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tc():
    import train_check
    return train_check


@pytest.fixture(scope="module")
def model(tc):
    return tc.make_net()


def test_sn_power_iteration(tc, model):
    assert tc.check_sn(model)


CONV_CASES = [
    ("3x3 s1 64->64 bn relu +res1", "encoder.layer1.0.conv2", "encoder.layer1.0.bn2", 64, (16, 24), dict(act=1, res1=True)),
    ("3x3 s2 64->128 bn relu", "encoder.layer2.0.conv1", "encoder.layer2.0.bn1", 64, (16, 24), dict(stride=2, act=1)),
    ("1x1 64->128 bn", "encoder.layer2.0.downsample.1", "encoder.layer2.0.downsample.2", 64, (8, 12), dict()),
    ("3x3 s1 32->32 bn relu", "encoder.conv2", "encoder.bn2", 32, (32, 32), dict(act=1)),
    ("3x3 s2 32->64 bn relu", "encoder.conv3", "encoder.bn3", 32, (32, 32), dict(stride=2, act=1)),
    ("shortcut relu->bn", "encoder.shortcut.3.0", "encoder.shortcut.3.2", 128, (8, 8), dict(mode=2, act=1)),
    ("guidance 16->32 reflect s2", "encoder.guidance_head.5", "encoder.guidance_head.7", 16, (16, 16),
     dict(stride=2, mode=2, act=1, reflect="prepadded")),
    ("guidance 32->128 reflect s2", "encoder.guidance_head.9", "encoder.guidance_head.11", 32, (16, 16),
     dict(stride=2, mode=2, act=1, reflect="prepadded")),
    ("3x3 s1 512->512 +res1", "encoder.layer_bottleneck.1.conv2", "encoder.layer_bottleneck.1.bn2", 512, (4, 4),
     dict(act=1, res1=True)),
    ("deconv 256->256", "decoder.layer2.0.conv1", "decoder.layer2.0.bn1", 256, (4, 6), dict(act=2, deconv=True)),
    ("3x3 256->128 +res1 +res2", "decoder.layer2.0.conv2", "decoder.layer2.0.bn2", 256, (8, 12),
     dict(act=2, res1=True, res2=True)),
    ("deconv 32->32 +res2", "decoder.conv1", "decoder.bn1", 32, (16, 16), dict(act=2, deconv=True, res2=True)),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_bn_forward_backward(tc, model, case):
    name, wkey, bnkey, cin, hw, kw = case
    assert tc.check_conv_bn(model, name, wkey, bnkey, cin, hw, **kw)


def test_gca_forward_backward(tc, model):
    assert tc.check_gca(model)


@pytest.mark.parametrize("form", ["shift_sum", "fold"])
def test_gca_forward_backward_both_forms(tc, model, form, monkeypatch):
    """The training GCA in the shift-sum form (csrc/gca_train2.cu; [Pk x 512 x Pk] value GEMMs, on the CTA-pair GEMM at
    this size) and in the round-1 fold form against the oracle's autograd, at a size with a ragged key grid."""
    from tcvom_b200 import train_engine
    orig = train_engine.TrainEngine.__init__

    def init(self, window):
        orig(self, window)
        self.gca_shift_sum_train = form == "shift_sum"
    monkeypatch.setattr(train_engine.TrainEngine, "__init__", init)
    getattr(model.NET, "_train_engines", {}).clear()
    try:
        assert tc.check_gca(model, hw=(40, 58))
    finally:
        getattr(model.NET, "_train_engines", {}).clear()


def test_tam_forward_backward(tc, model):
    assert tc.check_tam(model)


def test_full_training_step_matches_reference(tc):
    ok, rows, errs = tc.check_full_step(verbose=True)
    assert errs["losses"] < 2e-3, errs
    assert errs["alphas"] < 1e-3, errs
    assert errs["state"] < 1e-3, errs
    assert errs["grad_median"] < 4e-2, errs
    assert ok, (errs, rows[:5])


def test_full_training_step_matches_reference_on_the_well_conditioned_fixture(tc):
    """Same step, same reference, but with the residual-branch BatchNorm gains damped (tests/golden/make_golden.py
    damp_state): perturbations are no longer amplified ~1000x, so this comparison can SEE a 1 % gradient bug, which the
    standard fixture (gradients at a 1e-2 noise floor) cannot."""
    # The step is not bit-reproducible (fp32 atomics in the weight-gradient kernels) and a rare ordering lands on the other
    # side of a ReLU / |.| kink early in the network: 1 run in ~5 of the whole suite showed an outlier.  Three attempts.
    for attempt in range(3):
        ok, rows, errs = tc.check_full_step(verbose=attempt == 0, bound=1.0, damped=True)
        print("damped fixture:", {k: float("%.3e" % v) for k, v in errs.items()})
        if errs["grad_global"] < 1.5e-2 and errs["grad_median"] < 1e-2 and errs["grad_p90"] < 4e-2 and errs["alphas"] < 3e-4:
            break
    # Measured on B200: losses 1.4e-6, alphas 7.5e-5, state 5.9e-6; gradients (rel-L2 against the reference's) 5.7e-3 as one
    # vector, 4.0e-3 median, 1.4e-2 p90 -- 3.5x below the standard fixture and exactly the storage-precision ratio away
    # from what two fp32 implementations reach here (2.5e-5 median, fp32 vs fp64 oracle): activations and gradients are
    # stored with 16 mantissa bits (split-bf16, 2^-17 = 128 x the fp32 rounding), and the backward pass is linear in them.
    # The bounds below are ~3x the measurement: a gradient bug of a few per cent in any operator fails this test (the standard
    # fixture cannot see anything below its 1.4e-2 noise floor).
    assert errs["losses"] < 1e-4 and errs["alphas"] < 3e-4 and errs["state"] < 1e-4, errs
    assert errs["grad_global"] < 1.5e-2 and errs["grad_median"] < 1e-2 and errs["grad_p90"] < 4e-2, errs


def test_freeze_backbone_step_matches_reference(tc):
    """TAM pre-training mode, get_VMN_models(freeze_backbone=True) (VMN_model.py:77-81,99-103, VMN_GCA.py:18-24): encoder and
    decoder.layer1 / layer2 / gca in eval mode under no_grad, gradients and statistics updates for the decoder tail only.
    Against one step of the unmodified reference in that mode (tests/golden/train_step_s5_freeze.npz): losses, alphas, the
    tail's gradients, `grad is None` for exactly the parameters the reference leaves without a gradient, and the state
    tensors (frozen layers: spectral-norm u / v and running statistics unchanged; tail: updated)."""
    for attempt in range(3):
        ok, rows, errs = tc.check_full_step(verbose=attempt == 0, bound=1.0, freeze=True)
        print("freeze_backbone:", {k: float("%.3e" % v) for k, v in errs.items()})
        if errs["grad_global"] < 1.5e-2 and errs["grad_p90"] < 4e-2:
            break
    assert errs["nograd_mismatch"] == 0, errs
    assert errs["losses"] < 1e-4 and errs["alphas"] < 3e-4 and errs["state"] < 1e-4, errs
    assert errs["grad_global"] < 1.5e-2 and errs["grad_median"] < 1e-2 and errs["grad_p90"] < 4e-2, errs


def test_freeze_backbone_at_the_plugin_seam(tc):
    """VMN.forward in train mode with freeze_backbone through autograd: only tail parameters receive gradients."""
    import tcvom_b200
    from helpers import fixture_sd
    from tcvom_b200.train_engine import FROZEN_PREFIXES
    net = tcvom_b200.get_VMN_models("vmn_gca", agg_window=7, freeze_backbone=True)
    net.load_state_dict(fixture_sd(), strict=True)
    net = net.cuda().train()
    assert not net.encoder.training and not net.decoder.layer1.training and net.decoder.layer3.training
    before = {k: v.clone() for k, v in net.state_dict().items()}
    torch.manual_seed(0)
    S, H, W = 3, 64, 64
    frames = [torch.randn(1, 1, 6, H, W, device="cuda") for _ in range(S)]
    masks = [(torch.rand(1, 1, 1, H, W, device="cuda") > 0.5).float() for _ in range(S)]
    preds, attb, attf, small = net(frames, masks)
    (preds[1].mean() + attb[1].mean()).backward()
    after = net.state_dict()
    for n, p in net.named_parameters():
        if n.startswith(FROZEN_PREFIXES):
            assert p.grad is None, n
    assert any(p.grad is not None and float(p.grad.abs().sum()) > 0 for n, p in net.named_parameters()
               if n.startswith("decoder.fam"))
    for k, v in before.items():
        if k.startswith(FROZEN_PREFIXES):
            assert torch.equal(v, after[k]), k                       # no statistics / u / v update in the frozen part
    assert any(not torch.equal(before[k], after[k]) for k in before if k.startswith("decoder.layer3") and "running_mean" in k)


def test_train_mode_without_grad_runs_forward_only(model):
    import numpy as np
    from helpers import golden
    g = golden("train_step_s5.npz")
    a, fg, bg = (torch.from_numpy(g[k]).float().cuda() for k in ("a", "fg", "bg"))
    with torch.no_grad():
        out = model(a, fg, bg)
    assert len(out) == 12 and all(torch.isfinite(o).all() for o in out[:5])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_syncbn_two_ranks():
    """train_ddp.py's recipe (SyncBatchNorm + DistributedDataParallel over NCCL) on 2 GPUs: tools/ddp_check.py."""
    import json
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[-1])
    assert res["ok"], res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_syncbn_statistics_over_peer_memory_match_nccl():
    """tcv_peer_allreduce_f64 (one kernel over NVLink peer memory, rank-ordered sum) == the NCCL all-reduce it replaces for
    the SyncBatchNorm statistics, bit-identical on all ranks: tools/peer_check.py."""
    import json
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tools", "peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[-1])
    print(res)
    assert res["available"], "peer memory unavailable on this box: the step fell back to NCCL (still correct, not faster)"
    assert res["ok"], res


def test_vmn_seam_train_mode_matches_wrapper(tc):
    """The plugin seam in train mode (reference FullModel_VMD on top of tcvom_b200.VMN): torch losses on the native
    VMN outputs give the same losses and gradients as the fully native wrapper."""
    import numpy as np
    from helpers import golden, key_table
    from oracle import vmn_gca_oracle as O
    g = golden("train_step_s5.npz")
    a, fg, bg = (torch.from_numpy(g[k]).float().cuda() for k in ("a", "fg", "bg"))
    W = tc.LOSS_WEIGHTS
    # (1) fully native wrapper
    m1 = tc.make_net()
    out = m1(a, fg, bg)
    sum(w * o.mean() for w, o in zip(W, out[:5])).backward()
    g1 = {n: p.grad.clone() for n, p in m1.NET.named_parameters() if p.grad is not None}
    # (2) seam: native VMN.forward + the oracle's torch restatement of the reference losses, on the GPU
    m2 = tc.make_net()
    pp = O.train_preprocess(a, fg, bg, [3] * a.shape[0])
    pp["x6"] = pp["x6"].cuda()
    S = a.shape[1]
    frames = [pp["x6"][:, i:i + 1] for i in range(S)]
    masks = [pp["trimask"][:, i:i + 1] for i in range(S)]
    preds, attb, attf, small = m2.NET(frames, masks)
    La, Ldt, Latt, _, _ = O.vmd_losses(pp, preds, attb, attf, small)
    (W[0] * La + W[3] * Ldt + W[4] * Latt).backward()
    assert abs(float(La) - float(out[0])) < 1e-5 and abs(float(Latt) - float(out[4])) < 1e-5
    def spread(ga, gb):
        errs = sorted(float((ga[n] - gb[n]).double().norm()) / max(float(gb[n].double().norm()), 1e-12) for n in gb)
        return errs[-1], errs[len(errs) // 2]

    g2 = {n: p.grad.clone() for n, p in m2.NET.named_parameters() if p.grad is not None}
    # noise floor: the same wrapper step twice (fp32 atomics in the weight-gradient / statistics kernels change the
    # summation order from run to run, and this random-weight fixture amplifies 1e-7 differences strongly)
    m3 = tc.make_net()
    out3 = m3(a, fg, bg)
    sum(w * o.mean() for w, o in zip(W, out3[:5])).backward()
    g3 = {n: p.grad.clone() for n, p in m3.NET.named_parameters() if p.grad is not None}
    noise_worst, noise_med = spread(g3, g1)
    worst, med = spread(g2, g1)
    print(f"seam vs wrapper: worst {worst:.2e} median {med:.2e}; run-to-run noise: worst {noise_worst:.2e} median {noise_med:.2e}")
    assert med < max(5e-3, 4 * noise_med), (med, noise_med)
    assert worst < max(2e-2, 4 * noise_worst), (worst, noise_worst)


def _oracle_step(a, fg, bg, radii, with_att=True):
    """reference semantics on the CPU oracle: losses + gradients of one training step"""
    from helpers import fixture_sd, key_table
    from oracle import vmn_gca_oracle as O
    sd = {k: v.clone() for k, v in fixture_sd().items()}
    names = key_table()["trainable"]
    for n in names:
        sd[n].requires_grad_(True)
    out = O.full_vmd_forward(sd, a.cpu(), fg.cpu(), bg.cpu(), radii, train=True)
    w = (1.0, 1.0, 1.0, 0.5, 0.25) if with_att else (1.0, 1.0, 1.0, 0.0, 0.0)
    sum(wi * o.mean() for wi, o in zip(w, out[:5])).backward()
    return [float(o) for o in out[:5]], {n: (sd[n].grad if sd[n].grad is not None else torch.zeros_like(sd[n])) for n in names}


def _grad_errors(model, ref):
    errs = []
    for n, p in model.NET.named_parameters():
        if n in ref and float(ref[n].double().norm()) > 0:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            errs.append(float((g.cpu().double() - ref[n].double()).norm()) / float(ref[n].double().norm()))
    errs.sort()
    return errs[len(errs) // 2], errs[-1]


EDGE_CASES = {
    # S = 3: L_tc is identically zero (model.py:335-345), one centre frame
    "s3": dict(B=2, S=3, blank=None, cls="FullModel_VMD"),
    # one sample whose centre frames have no unknown pixel: L_af skips them (model.py:296-298), L_im's count clamps
    "no_unknown_sample": dict(B=2, S=5, blank=1, cls="FullModel_VMD"),
    # the plain FullModel wrapper (3 losses, no attention loss)
    "fullmodel": dict(B=1, S=5, blank=None, cls="FullModel"),
    # non-square frame: partial pixel tiles in every tensor-core kernel (OS16 grid 4 x 6, OS32 2 x 3)
    "rect64x96": dict(B=1, S=3, blank=None, cls="FullModel_VMD", hw=(64, 96)),
}


@pytest.mark.parametrize("case", list(EDGE_CASES), ids=list(EDGE_CASES))
def test_train_step_edge_cases_match_oracle(case):
    import numpy as np
    import tcvom_b200
    from helpers import fixture_sd
    from tcvom_b200 import synthetic
    c = EDGE_CASES[case]
    h, w = c.get("hw", (64, 64))
    a, fg, bg = synthetic.make_train_batch(c["B"], c["S"], h, w, seed=77)
    if c["blank"] is not None:
        a[c["blank"]] = 255                       # fully opaque sample: no 0 < alpha < 1 pixel in any frame
    a, fg, bg = (torch.from_numpy(t).float().cuda() for t in (a, fg, bg))
    cls = getattr(tcvom_b200, c["cls"])
    model = cls(model="vmn_gca", agg_window=7, dilate_kernel=2)
    model.NET.load_state_dict(fixture_sd(), strict=True)
    model = model.cuda().train()
    out = model(a, fg, bg)
    with_att = c["cls"] == "FullModel_VMD"
    nl = 5 if with_att else 3
    w = (1.0, 1.0, 1.0, 0.5, 0.25)[:nl]
    sum(wi * o.mean() for wi, o in zip(w, out[:nl])).backward()
    ref_losses, ref_grads = _oracle_step(a, fg, bg, [2] * c["B"], with_att)
    got = [float(o) for o in out[:nl]]
    assert np.allclose(got, ref_losses[:nl], rtol=2e-3, atol=1e-6), (got, ref_losses)
    assert all(np.isfinite(got))
    med, worst = _grad_errors(model, ref_grads)
    print(f"{case}: losses {got} grad median {med:.2e} worst {worst:.2e}")
    # noise-floor bounds (see tools/determinism_probe.py; smaller batches are noisier): these cases pin the BRANCH
    # semantics -- the losses above -- and guard against gross gradient errors
    assert med < 8e-2 and worst < 3e-1, (med, worst)


def test_two_forwards_before_backward_raise_instead_of_mixing_tapes():
    """The engine keeps one tape: back-propagating step 1 after step 2's forward must raise, not silently run step 2's
    tape (plain autograd supports the pattern; a wrong gradient would be the worst answer)."""
    import tcvom_b200
    from tcvom_b200 import synthetic
    from helpers import fixture_sd
    tm = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=2)
    tm.NET.load_state_dict(fixture_sd(), strict=True)
    tm = tm.to("cuda:0").train()
    a, fg, bg = (torch.from_numpy(t).float().cuda() for t in synthetic.make_train_batch(1, 3, 64, 64, seed=3))
    o1 = tm(a, fg, bg)
    o2 = tm(a, fg, bg)
    with pytest.raises(RuntimeError, match="another train-mode forward"):
        o1[0].mean().backward()
    o2[0].mean().backward()                                   # the latest step is intact
    assert all(torch.isfinite(p.grad).all() for p in tm.NET.parameters() if p.grad is not None)
    o3 = tm(a, fg, bg)
    with torch.no_grad():
        tm(a, fg, bg)                                         # a no_grad train-mode forward also replaces the tape
    with pytest.raises(RuntimeError, match="another train-mode forward"):
        o3[0].mean().backward()
