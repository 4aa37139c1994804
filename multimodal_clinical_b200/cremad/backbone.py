"""ResNet18 feature producers for Crema-D (audio: 1-channel spectrogram, visual: T RGB frames).

Encoders are outside the fused path (BASELINE.json north_star: "left in PyTorch and treated as feature
producers").  Built on torchvision's ResNet so the parameter names (conv1, bn1, layer1..4) are the
standard ones; ``forward`` returns the (B[*T], 512, H, W) feature map that FusionNet pools.
"""
import torch.nn as nn
from torchvision.models.resnet import BasicBlock, ResNet


class _Encoder(ResNet):
    def __init__(self, modality):
        super().__init__(BasicBlock, [2, 2, 2, 2])
        self.modality = modality
        if modality == 'audio':
            self.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
            nn.init.kaiming_normal_(self.conv1.weight, mode='fan_out', nonlinearity='relu')
        elif modality != 'visual':
            raise NotImplementedError(f"Incorrect modality, should be audio or visual but got {modality}")
        del self.fc, self.avgpool

    def forward(self, x):
        if self.modality == 'visual':
            B, C, T, H, W = x.size()
            x = x.permute(0, 2, 1, 3, 4).contiguous().view(B * T, C, H, W)
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


def resnet18(modality, progress=True, **kwargs):
    return _Encoder(modality)
