"""TEST INFRASTRUCTURE - plain PyTorch fp32 restatement of the reference's U-Net++ segmenter
(/root/reference/src/preprocessing/segmentation/model.py:8-83): same module tree, hence the same state_dict keys and, for
the same seed, the same initial parameters and outputs as the reference module (pinned by oracle/make_golden_unet.py,
frozen in tests/golden/unetpp.json).  The CUDA engine (csrc/k_unet.cu) is compared against this on the CPU."""
from __future__ import annotations

import torch
import torch.nn as nn


def conv_block(cin: int, cout: int) -> nn.Module:
    """model.py:8-21: (Conv3x3 -> BatchNorm -> ReLU) x 2 under the attribute name `conv`"""
    m = nn.Module()
    m.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                           nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
    m.forward = lambda x: m.conv(x)
    return m


class NestedUNetRef(nn.Module):
    def __init__(self, num_labels: int = 1, input_channels: int = 3):
        super().__init__()
        f = [64, 128, 256, 512, 1024]
        # creation order = the reference's (model.py:34-61), so a seeded initialisation draws the same numbers
        self.conv0_0 = conv_block(input_channels, f[0]); self.pool0 = nn.MaxPool2d(2)
        self.conv1_0 = conv_block(f[0], f[1]); self.pool1 = nn.MaxPool2d(2)
        self.conv2_0 = conv_block(f[1], f[2]); self.pool2 = nn.MaxPool2d(2)
        self.conv3_0 = conv_block(f[2], f[3]); self.pool3 = nn.MaxPool2d(2)
        self.conv4_0 = conv_block(f[3], f[4])
        self.up1_0 = conv_block(f[0] + f[1], f[0]); self.up2_0 = conv_block(f[1] + f[2], f[1]); self.up3_0 = conv_block(f[2] + f[3], f[2])
        self.up1_1 = conv_block(f[0] * 2 + f[1], f[0]); self.up2_1 = conv_block(f[1] * 2 + f[2], f[1])
        self.up1_2 = conv_block(f[0] * 3 + f[1], f[0])
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.final = nn.Conv2d(f[0], num_labels, kernel_size=1)

    def forward(self, x):
        up, cat = self.up, torch.cat
        x0_0 = self.conv0_0.conv(x)
        x1_0 = self.conv1_0.conv(self.pool0(x0_0))
        x2_0 = self.conv2_0.conv(self.pool1(x1_0))
        x3_0 = self.conv3_0.conv(self.pool2(x2_0))
        # model.py:69 also evaluates conv4_0(pool3(x3_0)); nothing consumes it
        x0_1 = self.up1_0.conv(cat([x0_0, up(x1_0)], 1))
        x1_1 = self.up2_0.conv(cat([x1_0, up(x2_0)], 1))
        x2_1 = self.up3_0.conv(cat([x2_0, up(x3_0)], 1))
        x0_2 = self.up1_1.conv(cat([x0_0, x0_1, up(x1_1)], 1))
        x1_2 = self.up2_1.conv(cat([x1_0, x1_1, up(x2_1)], 1))
        x0_3 = self.up1_2.conv(cat([x0_0, x0_1, x0_2, up(x1_2)], 1))
        return self.final(x0_3)


def seeded_model(seed: int = 0, randomize_bn: bool = True) -> NestedUNetRef:
    """Random-init model (there is no checkpoint in the reference repo) with non-trivial BatchNorm statistics so that the
    folded eval-mode BatchNorm is really exercised."""
    torch.manual_seed(seed)
    m = NestedUNetRef()
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=g)
                mod.bias.data = 0.2 * torch.randn(mod.bias.shape, generator=g)
                mod.running_mean = 0.1 * torch.randn(mod.running_mean.shape, generator=g)
                mod.running_var = 0.5 + torch.rand(mod.running_var.shape, generator=g)
    return m.eval()


def seeded_input(seed: int, n: int, h: int, w: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand((n, 1, h, w), generator=g).repeat(1, 3, 1, 1)        # grey / 255 replicated (inference.py:91-92)
