import sys, torch
sys.path.insert(0, '/root/repo')
from tests.test_task_gpu import _task, _batch
from tests.util import load_golden, rel_l2
g = load_golden("golden_1rank.npz")
task = _task(g)
enc = task.context_encoder
tokens = _batch(g)["contexts_ids"]
probe = torch.randn(8, 128, generator=torch.Generator().manual_seed(3)).cuda()
def run(lean):
    enc.lean_activations = lean
    enc.zero_grad()
    r = enc(tokens)
    (r * probe).sum().backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in enc.named_parameters() if p.grad is not None}
a = run(False); b = run(True); c = run(False)
for k in a:
    ra, rc = rel_l2(b[k], a[k]), rel_l2(c[k], a[k])
    if ra > 1e-3 or rc > 1e-3:
        print(f"{k:60s} lean-vs-full {ra:.3e}  full-vs-full {rc:.3e}  norm {float(a[k].norm()):.3e}")
print("done")
