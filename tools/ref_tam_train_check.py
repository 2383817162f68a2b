"""The reference's OWN vmn_index / vmn_dim networks (baseline/_ref, unmodified) training with the native TAM operator:
tcvom_b200.install(native_tam=True) swaps the FeatureAggregationModule class the reference decoders instantiate.  One
train-mode forward + backward of the same network with the reference TAM and with the native one (same weights, same
inputs, fp32, TF32 off): predictions, TAM logits and every gradient are compared.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from baseline import ref_env

ref_env.activate(cpu=False)
arch = sys.argv[1] if len(sys.argv) > 1 else "vmn_index"
import models.VMN as ref_vmn                       # the reference package
from models.VMN.VMN_model import FeatureAggregationModule as RefTAM

torch.manual_seed(0)
net_ref = ref_vmn.get_VMN_models(arch, agg_window=7).cuda()
net_ref.train(True)                                # the reference's VMN.train returns None (VMN_model.py:77-81)
import tcvom_b200

tcvom_b200.install(native_tam=True)
import models.VMN.VMN_Index as vi

assert vi.FeatureAggregationModule is tcvom_b200.FeatureAggregationModule
net_nat = ref_vmn.get_VMN_models._reference(arch, agg_window=7)       # the reference's factory, now building the native TAM
assert type(net_nat.decoder.fam) is tcvom_b200.FeatureAggregationModule and type(net_ref.decoder.fam) is RefTAM
net_nat.load_state_dict(net_ref.state_dict(), strict=True)
net_nat = net_nat.cuda()
net_nat.train(True)
S, H, W = 3, 64, 64
cin = 4
B = 2                                              # IndexNet's ASPP pools to 1x1 before a BatchNorm: batch > 1 in train mode
frames = [torch.randn(B, 1, cin, H, W, device="cuda") for _ in range(S)]
masks = [(torch.rand(B, 1, 1, H, W, device="cuda") > 0.4).float() for _ in range(S)]
def split_round(t):
    hi = t.to(torch.bfloat16).float()
    return hi + (t - hi).to(torch.bfloat16).float()


# noise floor: the reference TAM with its inputs and its output rounded to the 16 mantissa bits the native operator stores
# activations with (a straight-through rounding: gradients pass unchanged).  Whatever this arm differs from the plain
# reference by is the network amplifying a 2^-17 perturbation, not a difference between the operators.
class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return split_round(t)

    @staticmethod
    def backward(ctx, g):
        return g


hooks = []
res = {}
for name, net in (("ref", net_ref), ("nat", net_nat), ("floor", net_ref)):
    if name == "floor":
        fam = net.decoder.fam
        hooks.append(fam.register_forward_pre_hook(lambda m, a: tuple(_Round.apply(t) if i < 3 else t for i, t in enumerate(a))))
        hooks.append(fam.register_forward_hook(lambda m, a, o: (_Round.apply(o[0]),) + tuple(o[1:])))
    torch.manual_seed(7)                           # IndexNet's ASPP has a Dropout(0.5): same masks in both arms
    preds, attb, attf, small = net([f.clone() for f in frames], [m.clone() for m in masks])
    loss = preds[1].mean() + 0.1 * (attb[1] ** 2).mean() + 0.1 * attf[1].mean()
    net.zero_grad()
    loss.backward()
    res[name] = dict(pred=preds[1].detach(), attb=attb[1].detach(), small=small[1],
                     grads={n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None})
for h in hooks:
    h.remove()
rel = lambda a, b: float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-20))
g_ref, g_nat = res["ref"]["grads"], res["nat"]["grads"]
assert set(g_ref) == set(g_nat), set(g_ref) ^ set(g_nat)
gmax = max(float(g.norm()) for g in g_ref.values())
# conv biases in front of a batch-statistics BatchNorm have an exactly-zero true gradient: rounding noise on both sides
errs = sorted((rel(g_nat[n], g_ref[n]), n) for n in g_ref if float(g_ref[n].norm()) > 1e-6 * gmax)
num = sum(float((g_nat[n].double() - g_ref[n].double()).norm()) ** 2 for n in g_ref) ** 0.5
den = sum(float(g_ref[n].double().norm()) ** 2 for n in g_ref) ** 0.5
g_fl = res["floor"]["grads"]
fnum = sum(float((g_fl[n].double() - g_ref[n].double()).norm()) ** 2 for n in g_ref) ** 0.5
floor = dict(pred=rel(res["floor"]["pred"], res["ref"]["pred"]), grad_global=fnum / den,
             tam_value_weight=rel(g_fl["decoder.fam.value_conv.weight"], g_ref["decoder.fam.value_conv.weight"])
             if "decoder.fam.value_conv.weight" in g_ref else None)
out = dict(arch=arch, floor=floor, pred=rel(res["nat"]["pred"], res["ref"]["pred"]), attb=rel(res["nat"]["attb"], res["ref"]["attb"]),
           small_equal=bool(torch.equal(res["nat"]["small"], res["ref"]["small"])), n_grads=len(errs),
           grad_global=num / den, grad_median=errs[len(errs) // 2][0], grad_worst=errs[-1][0], worst_name=errs[-1][1],
           tam_grads={n: e for e, n in errs if n.startswith("decoder.fam")})
print(json.dumps(out))
