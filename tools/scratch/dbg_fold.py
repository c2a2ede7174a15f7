import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softspoken_b200 import checkpoint
from softspoken_b200._lib import lib, check
from softspoken_b200.engine import Engine
head = json.load(open(os.path.join(ROOT, "tests", "golden", "head_seed0.json")))
eng = Engine(checkpoint.synthetic_state_dict(0, head), 0, max_batch=4, mode="f16x3")
def dump(which, n):
    c, h, w = C.c_int(), C.c_int(), C.c_int()
    check(lib.ss_debug_activation(eng._ctx, which, n, None, C.byref(c), C.byref(h), C.byref(w), None))
    out = torch.empty((n, c.value, h.value, w.value), dtype=torch.float32, device="cuda")
    check(lib.ss_debug_activation(eng._ctx, which, n, C.c_void_p(out.data_ptr()), C.byref(c), C.byref(h), C.byref(w), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu()
torch.manual_seed(0)
n = 2
mel = torch.rand(n, 128, 256, device="cuda") * 1.5
os.environ["SS_TC_POOL_FOLD"] = "0"
a = eng.classify(mel); a0 = dump(0, n); a1 = dump(1, n); a2 = dump(2, n); ap = dump(12, n); aph = dump(12 + 0x100, n); apl = dump(12 + 0x200, n)
os.environ["SS_TC_POOL_FOLD"] = sys.argv[1] if len(sys.argv) > 1 else "3"
b = eng.classify(mel); b0 = dump(0, n); b1 = dump(1, n); b2 = dump(2, n); bp = dump(12, n); bph = dump(12 + 0x100, n); bpl = dump(12 + 0x200, n)
print('p1 hi differ', int((aph != bph).sum()), 'lo differ', int((apl != bpl).sum()), 'of', aph.numel())
dd = (apl != bpl).nonzero()
for i in dd[:8]:
    i = tuple(i.tolist()); print(i, 'hi', float(aph[i]) / 2, float(bph[i]) / 2, 'lo', float(apl[i]) / 2, float(bpl[i]) / 2, 'window', a0[i[0], i[1], 2*i[2]:2*i[2]+2, 2*i[3]:2*i[3]+2].flatten().tolist())
mp = torch.nn.functional.max_pool2d(a0, 2)
print('plain p1 vs maxpool(conv1):', int((ap != mp).sum()), ' folded p1 vs maxpool:', int((bp != mp).sum()), ' plain vs folded', int((ap != bp).sum()))
d = (bp != mp).nonzero()
if len(d):
    for i in d[:6]:
        i = tuple(i.tolist()); print(i, float(bp[i]), float(mp[i]), float(ap[i]), a0[i[0], i[1], 2*i[2]:2*i[2]+2, 2*i[3]:2*i[3]+2].flatten().tolist())
for name, x, y in (("conv1", a0, b0), ("conv2", a1, b1), ("conv3", a2, b2)):
    d = (x != y)
    print(name, "differing", int(d.sum()), "of", d.numel(), "max abs", float((x - y).abs().max()))
    if d.any():
        idx = d.nonzero()
        print(" first", idx[:8].tolist(), " last", idx[-4:].tolist())
        print(" by image", d.sum(dim=(1, 2, 3)).tolist())
        print(" rows with diffs", torch.unique(idx[:, 2])[:20].tolist(), "cols", torch.unique(idx[:, 3])[:20].tolist(), "ch", torch.unique(idx[:, 1])[:40].tolist())
        i = idx[0]; print(" vals", float(x[tuple(i)]), float(y[tuple(i)]))
print("logits equal", torch.equal(a, b), float((a - b).abs().max()))
