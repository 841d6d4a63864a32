"""Segmentation head (drop-in for Dino/modules/segmentor.py:37-95, same parameter names / shapes).

ROUND-1 STATUS: the convolutions / transposed convolutions / BatchNorm of this head still run through PyTorch
(cuDNN, channels-last, bf16 autocast for the convs) -- they are ~8 % of the step FLOPs and are the
one part of the hot path that is NOT yet hand-written CUDA.  Its inputs (the three norm_seg taps) and its loss
(ccd_seg_ce_fwd/bwd) are on the sm_100a kernels.  `conv_mla` is constructed but never called, exactly like the
reference (its parameters never receive gradients: train.py:106 find_unused_parameters=True).
"""
import torch
import torch.nn as nn


def _cbr(cin, cout, k, pad):
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class Conv_MLA(nn.Module):          # segmentor.py:6-35 -- parameters only (unused by forward)
    def __init__(self, in_channels=1024, mla_channels=256):
        super().__init__()
        self.mla_p2_1x1 = _cbr(in_channels, mla_channels, 1, 0)
        self.mla_p3_1x1 = _cbr(in_channels, mla_channels, 1, 0)
        self.mla_p4_1x1 = _cbr(in_channels, mla_channels, 1, 0)
        self.mla_p2 = _cbr(mla_channels, mla_channels, 3, 1)
        self.mla_p3 = _cbr(mla_channels, mla_channels, 3, 1)
        self.mla_p4 = _cbr(mla_channels, mla_channels, 3, 1)


class MLAHead(nn.Module):           # segmentor.py:37-70
    def __init__(self, in_channels=384, mla_channels=128, mlahead_channels=64):
        super().__init__()

        def branch():
            return nn.Sequential(nn.Conv2d(in_channels, mla_channels, 3, padding=1, bias=False), nn.BatchNorm2d(mla_channels),
                                 nn.ReLU(), nn.Conv2d(mla_channels, mlahead_channels, 1, bias=False),
                                 nn.BatchNorm2d(mlahead_channels), nn.ReLU())
        self.head2, self.head3, self.head4 = branch(), branch(), branch()

    def forward(self, p2, p3, p4):
        return torch.cat([self.head2(p2), self.head3(p3), self.head4(p4)], dim=1)


class SegHead(nn.Module):           # segmentor.py:73-95
    def __init__(self, in_channels=384, mla_channels=128, mlahead_channels=64, num_classes=2, **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.conv_mla = Conv_MLA(in_channels, mla_channels)
        self.mlahead = MLAHead(in_channels=in_channels, mla_channels=mla_channels, mlahead_channels=mlahead_channels)
        self.unpool1 = nn.Sequential(nn.ConvTranspose2d(192, 128, (4, 4), (2, 2), (1, 1)), nn.BatchNorm2d(128), nn.ReLU(True))
        self.unpool2 = nn.Sequential(nn.ConvTranspose2d(128, 128, (4, 4), (2, 2), (1, 1)), nn.BatchNorm2d(128), nn.ReLU(True))
        self.cls = nn.Conv2d(128, self.num_classes, 3, padding=1)
        self.autocast_bf16 = True

    def forward(self, inputs):
        dev = inputs[0].device.type
        with torch.autocast(device_type=dev, dtype=torch.bfloat16, enabled=self.autocast_bf16 and dev == "cuda"):
            x = self.mlahead(inputs[0], inputs[1], inputs[2])
            x = self.unpool1(x)
            x = self.unpool2(x)
            x = self.cls(x)
        return x.float()
