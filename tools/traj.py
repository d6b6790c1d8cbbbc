"""Loss trajectory of a short training run (race hunting: must be identical across settings)."""
import sys, os, numpy as np, torch, random
sys.path.insert(0, "/root/repo")
from dincae_b200 import synthetic
from dincae_b200.data import DeviceCube
from dincae_b200.engine import Plan
from dincae_b200.sampler import TrainSampler
from dincae_b200.trainer import Trainer
use_graph = "--graph" in sys.argv
lon, lat, time_, data, mask = synthetic.sst_like_cube(T=int(os.environ.get('TRAJ_T','300')), H=112, W=112, seed=12345)
missing = np.ma.getmaskarray(data)
cube = DeviceCube(lon, lat, time_, [data], [missing], [1.0], 3)
plan = Plan(112, 112, cube.nvar)
tr = Trainer(plan, cube, 50, noise_seed=1234, jitter_std=[0.05], use_graph=use_graph)
tr.set_lr(1e-3)
random.seed(99)
sm = TrainSampler(int(os.environ.get('TRAJ_T','300')), 50, 45, seed=7)
out = []
for k in range(int(os.environ.get('TRAJ_N','12'))):
    tr.train_step(*sm.next_batch())
    out += tr.flush_losses()
print(" ".join("%.6f" % c for c, _ in out[-6:]))
