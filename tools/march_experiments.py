"""Times the marching 32->32 k3 conv at the frustum size with experiment switches (SS_MARCH_DBG bits:
1 no MMAs, 2 no stores, 4 no stats, 8 no fix-up math, 16 no plane TMA after the first ring fill)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for dbg in (0,):
        env = dict(os.environ, SS_MARCH_DBG=str(dbg))
        subprocess.run([sys.executable, __file__, str(dbg)], env=env)
    sys.exit(0)
import torch
from stereoscene_b200 import ops
from stereoscene_b200.ops import Vol
dev = "cuda"
torch.manual_seed(0)
conv = torch.nn.Conv3d(32, 32, 3, 1, 1, bias=False).to(dev)
gn = torch.nn.GroupNorm(2, 32).to(dev)
x = torch.randn(1, 112, 48, 160, 32, device=dev)
sc = torch.rand(1, 32, device=dev) + 0.5
sh = torch.randn(1, 32, device=dev) * 0.1
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def graphed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    return lambda: g.replay(), reps
def tg(fn):
    r, reps = graphed(fn)
    return t(r, 5) / reps
plain = tg(lambda: ops.conv(Vol(x), conv, want_stats=True))
pend = tg(lambda: ops.conv(Vol(x, sc, sh, ops.SS_ACT_RELU), conv, want_stats=True))
nostat = tg(lambda: ops.conv(Vol(x), conv))
print(f"dbg={sys.argv[1]:>3}: plain+stats {plain:.4f} ms   pending+stats {pend:.4f} ms   plain no stats {nostat:.4f} ms", flush=True)
