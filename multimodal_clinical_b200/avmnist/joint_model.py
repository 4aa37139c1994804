"""AV-MNIST, image + spectrogram (avmnist/joint_model.py of the reference): two LeNet encoders whose pooled outputs have
DIFFERENT widths (48 and 192), mean fusion of the two ten-way heads.  The heads (``classifier_x{1,2}``, same state-dict
names) and the loss run on the fused multi-head kernel; the LeNets stay in PyTorch as feature producers."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..multi import FusedMeanFusionHeads
from ..utils.MultiModel import MeanFusionMultiBaseModel


class GlobalPooling2D(nn.Module):
    def forward(self, x):
        return x.flatten(2).mean(2)


class LeNet(nn.Module):
    """Conv-BN-ReLU-maxpool blocks with doubling channel counts (avmnist/joint_model.py:32-97; parameter names
    ``convs.i`` / ``bns.i`` as in the reference)."""

    def __init__(self, in_channels, args_channels, additional_layers, output_each_layer=False, linear=None, squeeze_output=True):
        super().__init__()
        self.output_each_layer = output_each_layer
        chans = [in_channels] + [args_channels * 2 ** i for i in range(additional_layers + 1)]
        self.convs = nn.ModuleList([nn.Conv2d(chans[0], chans[1], kernel_size=5, padding=2, bias=False)] +
                                   [nn.Conv2d(chans[i], chans[i + 1], kernel_size=3, padding=1, bias=False)
                                    for i in range(1, additional_layers + 1)])
        self.bns = nn.ModuleList([nn.BatchNorm2d(c) for c in chans[1:]])
        self.gps = nn.ModuleList([GlobalPooling2D() for _ in chans[1:]])
        self.sq_out = squeeze_output
        self.linear = nn.Linear(linear[0], linear[1]) if linear is not None else None
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_uniform_(m.weight)

    def forward(self, x):
        tempouts, out = [], x
        for conv, bn, gp in zip(self.convs, self.bns, self.gps):
            out = F.max_pool2d(F.relu(bn(conv(out))), 2)
            tempouts.append(gp(out))
        if self.linear is not None:
            out = self.linear(out)
        tempouts.append(out)
        if self.output_each_layer:
            return [t.squeeze() for t in tempouts] if self.sq_out else tempouts
        return out.squeeze() if self.sq_out else out


class FusionNet(nn.Module):
    def __init__(self, num_classes, loss_fn):
        super().__init__()
        self.x1_model = LeNet(1, 6, 3)
        self.x2_model = LeNet(1, 6, 5)
        self.classifier_x1 = nn.Linear(48, num_classes)
        self.classifier_x2 = nn.Linear(192, num_classes)
        self.num_classes = num_classes
        self.loss_fn = loss_fn
        self.fused = FusedMeanFusionHeads(num_classes)

    def forward(self, x1_data, x2_data, label):
        """-> (x1_logits, x2_logits, avg_logits, loss)   avmnist/joint_model.py:117-138"""
        f1 = F.relu(self.x1_model(x1_data)).reshape(x1_data.shape[0], -1)
        f2 = F.relu(self.x2_model(x2_data)).reshape(x2_data.shape[0], -1)
        return self.fused([f1, f2], [self.classifier_x1, self.classifier_x2], label)


class MultimodalAVMnistModel(MeanFusionMultiBaseModel):
    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss())
