"""Run a few refiner training steps (BASELINE config 5 shapes) -- target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoposeestimation_b200 import synthetic as synth
from autoposeestimation_b200.densefusion.train_refiner import RefinerTrainer
B, N, O = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 500, 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(4)
pts = torch.randn((B, N, 3), device=dev, generator=g) * 0.05
emb = torch.randn((B, 32, N), device=dev, generator=g)
idx = torch.randint(0, O, (B,), device=dev, generator=g)
model = (torch.rand((B, N, 3), device=dev, generator=g) - 0.5) * 0.2
target = model + 0.01
sd = synth.refiner_state_dict(1007, O)
sd['conv3_r.bias'] = sd['conv3_r.bias'].copy(); sd['conv3_r.bias'][0::4] += 1.0
tr = RefinerTrainer(sd, O, B, N, sym_list=[0])
for _ in range(steps):
    d = tr.train_step(pts, emb, idx, target, model)
torch.cuda.synchronize()
print('dis', float(d.mean()))
