import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eventgrad_b200.models import build_model
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
a = build_model("resnet18").cuda().train(); b = build_model("resnet18").cuda().train(); b.load_state_dict(a.state_dict())
x = torch.randn(32, 3, 32, 32, device="cuda"); y = torch.randint(0, 10, (32,), device="cuda")
acts = {0: {}, 1: {}}
def hook(tag, name):
    def f(m, i, o): acts[tag][name] = o.detach().clone()
    return f
for tag, m in ((0, a), (1, b)):
    for n, mod in m.named_modules():
        if mod.__class__.__name__ in ("FusedBNAct", "ShadowConv2d"): mod.register_forward_hook(hook(tag, n))
la = F.cross_entropy(a(x), y); la.backward()
os.environ["EGB_FUSED_BN"] = "0"
lb = F.cross_entropy(b(x), y); lb.backward()
print("loss", float(la), float(lb))
for n in acts[0]:
    u, v = acts[0][n], acts[1][n]
    print(f"fwd {n:40s} rel {float((u-v).norm()/(v.norm()+1e-12)):.3e}")
for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
    print(f"grad {n:40s} rel {float((p.grad-q.grad).norm()/(q.grad.norm()+1e-12)):.3e}  norm {float(q.grad.norm()):.3e}")
