import sys, cProfile, pstats
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, torch
from opticalflowfromdepth_b200 import inloop, synthetic
dev = torch.device("cuda:0")
H, W, B = 368, 496, 8
fr = [synthetic.diml_frame(200 + k % 16, H, W) for k in range(B)]
img = torch.from_numpy(np.stack([f[0] for f in fr])).to(dev)
dep = torch.from_numpy(np.stack([f[1] for f in fr])).to(dev)
s = inloop.InLoopSampler(dev, seed=4)
for _ in range(5):
    s(img, dep).raft_tuple()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    s(img, dep).raft_tuple()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
