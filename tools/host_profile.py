#!/usr/bin/env python
"""Host-side cost of one training step (GPU box): cProfile over a few bs=64 steps, top functions by cumulative time."""
import cProfile, pstats, os, sys, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
from mmlf_b200.model.feed_forward import FeedForward
from mmlf_b200.model import loss as L
from mmlf_b200.optim import FusedAdam
from mmlf_b200 import _lib
dev = 'cuda'
torch.manual_seed(0)
model = FeedForward(**bench.model_kwargs('base')).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
views = [torch.rand((B, 9, 3, 96, 96), device=dev) for _ in range(4)]
gt = torch.rand((B, 96, 96), device=dev)
mask = L.create_mask_margin((B, 96, 96), 11).to(torch.int32).to(dev)
opt = FusedAdam(model.parameters(), lr=1e-3)
fn = L.MaskedL1Loss()
model.train()
def step():
    opt.zero_grad()
    out = model(*views)
    l = fn(out, gt, mask)
    l.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5): step()
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
print(f'enqueue {1e3*(t1-t0)/5:.2f} ms/step, total {1e3*(t2-t0)/5:.2f} ms/step, launches/step {_lib.launch_count/8}')
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
print(s.getvalue()[:6000])
