"""GPU: the standalone operator modules in train mode -- tcvom_b200.FeatureAggregationModule (VMN_model.py:9-68) and
tcvom_b200.GuidedCxtAtten (GCA/ops.py:83-229) called with autograd, the way the reference's own networks call them
(VMN_DIM / VMN_Index / VMN_FBA decoders call `self.fam(...)` inside an ordinary autograd graph).  Checked against the CPU
oracle's torch restatement differentiated by torch autograd on the same seeded inputs: outputs, input gradients, parameter
gradients, BatchNorm running statistics."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-20))


def split_round(t):
    """the value the kernels see: fp32 stored as two bf16 planes (16 mantissa bits)"""
    hi = t.to(torch.bfloat16).float()
    return hi + (t - hi).to(torch.bfloat16).float()


@pytest.mark.parametrize("chn", [32, 128, 256])
def test_tam_operator_train_mode_matches_oracle_autograd(chn):
    import tcvom_b200
    from oracle import vmn_gca_oracle as O
    torch.manual_seed(chn)
    fam = tcvom_b200.FeatureAggregationModule(chn, 1, 7)
    sd = {k: torch.randn_like(v) * (0.3 / (chn * 9) ** 0.5 if k.endswith("weight") else 0.1) for k, v in fam.state_dict().items()}
    fam.load_state_dict(sd, strict=True)
    fam = fam.cuda().train()
    B, H, W = 2, 10, 14
    xs = [split_round(torch.randn(B, chn, H, W)) for _ in range(3)]
    mask = (torch.rand(B, 1, 8 * H, 8 * W) > 0.4).float()
    dy = split_round(torch.randn(B, chn, H, W))
    db, df = torch.randn(B, 49, H * W) * 0.1, torch.randn(B, 49, H * W) * 0.1
    # native, through autograd
    gx = [t.clone().cuda().requires_grad_(True) for t in xs]
    feat, attb, attf, sm = fam(gx[0], gx[1], gx[2], mask.cuda())
    assert feat.requires_grad and attb.requires_grad and not sm.requires_grad
    ((feat * dy.cuda()).sum() + (attb * db.cuda()).sum() + (attf * df.cuda()).sum()).backward()
    # oracle + torch autograd on the CPU
    rsd = {"fam." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rx = [t.clone().requires_grad_(True) for t in xs]
    rfeat, rb, rf, rm = O.tam(rsd, "fam", rx[0], rx[1], rx[2], mask, 7)
    ((rfeat * dy).sum() + (rb * db).sum() + (rf * df).sum()).backward()
    assert torch.equal(sm.cpu(), rm)
    errs = dict(feat=rel(feat, rfeat), attb=rel(attb, rb), attf=rel(attf, rf))
    for i, n in enumerate(("dx", "db", "df")):
        errs[n] = rel(gx[i].grad, rx[i].grad)
    for n, p in fam.named_parameters():
        errs["d" + n] = rel(p.grad, rsd["fam." + n].grad)
    print(chn, {k: float("%.2e" % v) for k, v in errs.items()})
    assert max(errs.values()) < 2e-3, errs
    # a second step accumulates into .grad like any autograd node
    g0 = fam.key_conv.weight.grad.clone()
    feat2, _, _, _ = fam(gx[0], gx[1], gx[2], mask.cuda())
    (feat2 * dy.cuda()).sum().backward()
    assert not torch.equal(fam.key_conv.weight.grad, g0)


def test_tam_operator_second_backward_raises():
    import tcvom_b200
    fam = tcvom_b200.FeatureAggregationModule(32, 1, 3).cuda().train()
    x = torch.randn(1, 32, 8, 8, device="cuda", requires_grad=True)
    m = torch.ones(1, 1, 64, 64, device="cuda")
    y1 = fam(x, x, x, m)[0]
    y2 = fam(x, x, x, m)[0]              # replaces the module's tape
    y2.sum().backward()
    with pytest.raises(RuntimeError):
        y1.sum().backward()


def test_gca_operator_train_mode_matches_oracle_autograd():
    import tcvom_b200
    from helpers import fixture_sd
    from oracle import vmn_gca_oracle as O
    torch.manual_seed(3)
    full = fixture_sd()
    sd = {k[len("decoder.gca."):]: v.clone() for k, v in full.items() if k.startswith("decoder.gca.")}
    sd["W.1.weight"] = torch.full_like(sd["W.1.weight"], 0.5)     # a visible attention branch (the init is 1e-3)
    gca = tcvom_b200.GuidedCxtAtten(128, 128)
    gca.load_state_dict(sd, strict=True)
    gca = gca.cuda().train()
    B, H, W = 2, 16, 24
    f = split_round(torch.randn(B, 128, H, W))
    al = split_round(torch.randn(B, 128, H, W))
    unk = (torch.rand(B, 1, H, W) > 0.3).float()
    dy = split_round(torch.randn(B, 128, H, W))
    gf, ga = f.clone().cuda().requires_grad_(True), al.clone().cuda().requires_grad_(True)
    y, (offsets, scale) = gca(gf, ga, unk.cuda())
    (y * dy.cuda()).sum().backward()
    rsd = {"g." + k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
           for k, v in sd.items()}
    rf, ra = f.clone().requires_grad_(True), al.clone().requires_grad_(True)
    with O.TrainMode():
        ry = O.gca_attention(rsd, "g", rf, ra, unk)
    (ry * dy).sum().backward()
    errs = dict(y=rel(y, ry), df=rel(gf.grad, rf.grad), dalpha=rel(ga.grad, ra.grad))
    for n, p in gca.named_parameters():
        errs["d" + n] = rel(p.grad, rsd["g." + n].grad)
    st = gca.state_dict()
    for k in ("W.1.running_mean", "W.1.running_var"):
        errs[k] = float((st[k].cpu() - rsd["g." + k]).abs().max())
    assert int(st["W.1.num_batches_tracked"]) == int(rsd["g.W.1.num_batches_tracked"])
    print({k: float("%.2e" % v) for k, v in errs.items()})
    assert max(errs.values()) < 3e-3, errs
    # train mode under no_grad: forward only, same values
    with torch.no_grad():
        y2, _ = gca(f.cuda(), al.cuda(), unk.cuda())
    assert float((y2 - y).abs().max()) < 1e-4
