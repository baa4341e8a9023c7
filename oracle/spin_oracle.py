"""CPU restatement of the SPIN / HMR ResNet-50 feature extractor ((f)2, the step BEFORE the path).  TEST INFRASTRUCTURE.

Functional (state_dict in, tensors out) restatement of reference lib/models/spin.py:
  * `Bottleneck.forward`        :38-57   conv1x1-BN-ReLU, conv3x3(stride)-BN-ReLU, conv1x1-BN, (+downsample), add, ReLU
  * `HMR._make_layer`           :113-127 first block of a layer carries the stride and a 1x1-conv + BN downsample
  * `HMR.feature_extractor`     :129-143 conv7x7/2-BN-ReLU-maxpool3x3/2, layer1-4 ([3,4,6,3] bottlenecks), avgpool 7 -> [B,2048]
BatchNorm runs in eval mode (running statistics), as in the demo's use (main/run_demo.py:249,315: `hmr().to(device)` +
`.eval()`); pinned against the reference module itself by oracle/gen_golden.py::gen_spin (fixture tests/golden/spin_B2.npz).
Only tests/ and bench.py's cpu legs may import this file."""
import torch
import torch.nn.functional as F

LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def bottleneck(sd, p, x, stride, has_down):
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    res = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride)) if has_down else x
    return F.relu(out + res)


def feature_extractor(sd, x, return_intermediates=False):
    """x [B,3,224,224] -> xf [B,2048]."""
    inter = {}
    x = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    inter["stem"] = x
    for li, (planes, blocks, stride) in enumerate(LAYERS, 1):
        for bi in range(blocks):
            x = bottleneck(sd, f"layer{li}.{bi}", x, stride if bi == 0 else 1, bi == 0)
        inter[f"layer{li}"] = x
    xf = F.avg_pool2d(x, 7, stride=1).reshape(x.shape[0], -1)
    return (xf, inter) if return_intermediates else xf
