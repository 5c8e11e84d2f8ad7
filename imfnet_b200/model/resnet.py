"""ResNet image backbone: the PARAMETER TREE of the reference's torchvision-style ResNet
(/root/reference/model/resnet.py:120-216: conv7x7/2 -> BN -> ReLU -> maxpool/2 -> layer1 -> layer2 are what runs; layer3/4/fc are
constructed so checkpoints load strictly, but never executed).  The arithmetic of the executed prefix lives in
Img_Encoder.ImagePlan (sm_100a kernels); these modules only hold weights under the reference's names."""
import torch
import torch.nn as nn

__all__ = ['ResNet', 'resnet18', 'resnet34']


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    def __init__(self, in_channels, block, layers, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._stage(block, 64, layers[0], 1)
        self.layer2 = self._stage(block, 128, layers[1], 2)
        self.layer3 = self._stage(block, 256, layers[2], 2)   # weights only (state_dict contract)
        self.layer4 = self._stage(block, 512, layers[3], 2)   # weights only
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def _stage(self, block, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                 nn.BatchNorm2d(planes * block.expansion))
        mods = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        mods += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    def forward(self, x):
        raise NotImplementedError("the backbone is a parameter container: run it through imfnet_b200.model.ImageEncoder "
                                  "(Img_Encoder.ImagePlan executes conv1..layer2 on the sm_100a kernels)")


def resnet18(in_channels=3, pretrained=False, progress=True, **kwargs):
    return ResNet(in_channels, BasicBlock, [2, 2, 2, 2], **kwargs)


def resnet34(in_channels=3, pretrained=False, progress=True, **kwargs):
    """`pretrained` is accepted for signature compatibility; weights always come from load_state_dict
    (the reference's ImageNet download, Img_Encoder.py:13, is overwritten by the checkpoint anyway)."""
    return ResNet(in_channels, BasicBlock, [3, 4, 6, 3], **kwargs)
