"""Where the one-off time of DINCAE.reconstruct_gridded_nc goes (cProfile, one epoch)."""
import cProfile
import os
import pstats
import sys
import tempfile

sys.path.insert(0, "/root/repo")
import DINCAE  # noqa: E402
from dincae_b200 import synthetic  # noqa: E402

d = tempfile.mkdtemp()
lon, lat, time_, data, mask = synthetic.sst_like_cube(T=5266, H=112, W=112, seed=12345)
fname = os.path.join(d, "sst.nc")
synthetic.write_input_file(fname, "SST", lon, lat, time_, data, mask)
import torch  # noqa: E402
torch.zeros(1).cuda()
pr = cProfile.Profile()
pr.enable()
DINCAE.reconstruct_gridded_nc(fname, "SST", os.path.join(d, "out"), epochs=1, batch_size=50, save_each=0,
                              save_model_each=0, iseed=1, loss=[])
pr.disable()
st = pstats.Stats(pr, stream=sys.stderr)
st.sort_stats("cumulative").print_stats(45)
st.print_callees("run_step")
st.print_callees("reconstruct$")
