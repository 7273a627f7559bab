"""GPU check awaiting a free B200 slot; moved into tests/test_gpu_net.py once it has passed there."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np, torch
from oracle.models import resunet_forward
from test_gpu_net import _randomise_bn, _psnr


def test_resunet_nine_input_channels():
    """More than 7 input channels (9*C > 64): the normalised input is an ordinary NHWC 3x3 source instead of an im2col block."""
    from pssr2_b200.models import ResUNet
    torch.manual_seed(2)
    model = ResUNet(channels=[9, 1], hidden=[64, 128], scale=4, depth=1).eval()
    _randomise_bn(model, 3)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.tensor(np.random.default_rng(2).integers(0, 256, (2, 9, 128, 128)).astype(np.float32))
    want = resunet_forward(sd, x)
    got = model.cuda()(x.cuda()).cpu()
    assert got.shape == want.shape == (2, 1, 512, 512)
    assert float((got - want).abs().max()) <= 3e-2 and _psnr(got, want) >= 50.0



if __name__ == "__main__":
    test_resunet_nine_input_channels()
    print("nine-channel ResUNet ok")
