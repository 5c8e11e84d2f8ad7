"""ImageEncoder().forward(x): [B,3,H,W] -> [B,128,H/8,W/8]  (/root/reference/model/Img_Encoder.py:9-18)."""
import torch.nn as nn

from . import resnet


class ImageEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = resnet.resnet34(in_channels=3, pretrained=False, progress=False)

    def forward(self, x):
        return self.backbone(x)
