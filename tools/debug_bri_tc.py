"""Debug aid: which of the two tcgen05 products of the BRI kernel disagrees with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from stereoscene_b200 import ops, cabi
cabi.load()
B, D, H, W = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 48, 8, 16))]
torch.manual_seed(10)
q = F.softmax(torch.randn(B, D, H, W) * 2, dim=1)
kv = F.softmax(torch.randn(B, D, H, W) * 2, dim=1)
N = H * W
wq, bq, wk, bk, wv, bv, gamma = 6.5, 0.1, 5.5, -0.05, 1.3, 0.02, 0.5
qf, kf = q.reshape(B, D, N).double(), kv.reshape(B, D, N).double()
conf = F.softmax(qf, dim=1).max(dim=1)[0]
S0 = torch.bmm(qf.transpose(1, 2), kf)
sk = kf.sum(1)
def model(s0_ok, o_ok):
    E = (wq * wk * S0 if s0_ok else 0 * S0) + bq * wk * sk[:, None, :]
    P = torch.exp(E - E.max(-1, keepdim=True)[0])
    l = P.sum(-1)
    Pc = P * conf[:, None, :]
    r = Pc.sum(-1)
    O0 = torch.bmm(kf, Pc.transpose(1, 2)) if o_ok else torch.zeros(B, D, N, dtype=torch.float64)
    return gamma * (wv * O0 + bv * r[:, None, :]) / l[:, None, :] + kf
params = torch.tensor([wq, bq, wk, bk, wv, bv, gamma]).cuda()
both = torch.zeros(B, D, H, W, 1, device="cuda")
ops.bri_attention(q.cuda(), kv.cuda(), params, both[..., 0], 1)
got = both[..., 0].reshape(B, D, N).double().cpu()
for s0_ok in (1, 0):
    for o_ok in (1, 0):
        w = model(s0_ok, o_ok)
        print(f"S0 {'ok' if s0_ok else '0 '} O {'ok' if o_ok else '0 '}: rel err {float((got - w).abs().max() / w.abs().max()):.3e}")
w = model(1, 1)
err = (got - w).abs()
print("err by query block of 32:", [f"{float(err[:, :, k:k+32].max()):.1e}" for k in range(0, min(N, 256), 32)])
print("err by depth block of 8:", [f"{float(err[:, k:k+8, :].max()):.1e}" for k in range(0, D, 8)])
