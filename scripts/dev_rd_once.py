import sys
sys.path.insert(0, ".")
import torch
from pssr2_b200.models import RDResUNet
torch.manual_seed(0)
m = RDResUNet().eval().cuda()
x = torch.randint(0, 256, (50, 1, 128, 128), device="cuda").float()
st, _ = m._state(x)
plan = st["plan"]
for _ in range(2): plan.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
