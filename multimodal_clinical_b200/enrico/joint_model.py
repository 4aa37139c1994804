"""Enrico late fusion of screenshot + wireframe (enrico/joint_model.py of the reference) on the fused step.
The classifier lives inside each ``x_model`` (state-dict keys ``x{1,2}_model.classifier.*``) and the
ResNet18 features are frozen, so the step runs without feature gradients."""
import torch
import torch.nn as nn
from torch.optim.lr_scheduler import StepLR
from torchvision import models as tmodels

from ..heads import FusedLateFusionHead
from ..utils.BaseModel import JointLogitsBaseModel


class ResNet18Slim(nn.Module):
    def __init__(self, hiddim, pretrained=True, freeze_features=True):
        super().__init__()
        self.hiddim = hiddim
        try:
            backbone = tmodels.resnet18(weights="DEFAULT" if pretrained else None)
        except Exception:                      # no network: random init (weights come from a checkpoint)
            backbone = tmodels.resnet18(weights=None)
        self.model = nn.Sequential(*list(backbone.children())[:-1])
        self.embedding = nn.AdaptiveAvgPool2d((1, 1))
        self.classifier = nn.Linear(512, hiddim)
        if freeze_features:
            for param in self.model.parameters():
                param.requires_grad = False

    def embed(self, x):
        features = self.model(x)
        return self.embedding(features).view(features.size(0), -1)

    def forward(self, x):
        embedding = self.embed(x)
        return embedding, self.classifier(embedding)


class FusionNet(nn.Module):
    def __init__(self, num_classes, loss_fn):
        super().__init__()
        self.x1_model = ResNet18Slim(num_classes)
        self.x2_model = ResNet18Slim(num_classes)
        self.num_classes = num_classes
        self.loss_fn = loss_fn
        self.fused = FusedLateFusionHead(num_classes, mode="jlogits")

    def forward(self, x1_data, x2_data, label):
        e1, e2 = self.x1_model.embed(x1_data), self.x2_model.embed(x2_data)
        return self.fused(e1, e2, self.x1_model.classifier, self.x2_model.classifier, label)


class MultimodalEnricoModel(JointLogitsBaseModel):
    def __init__(self, args):
        super().__init__(args)

    def configure_optimizers(self):
        optimizer = self._sgd()
        if self.args.use_scheduler:
            scheduler = {'scheduler': StepLR(optimizer, step_size=10, gamma=0.5), 'interval': 'epoch', 'frequency': 1}
            return [optimizer], [scheduler]
        return optimizer

    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss())
