"""Food101 plain late fusion (food101/joint_model.py of the reference): SigLIP embeddings -> per-modality
MLP -> mean of the logits -> CE, on the fused step (the MLPs' last Linear is the fused head's classifier)."""
import torch
import torch.nn as nn
from torch.optim.lr_scheduler import StepLR

from ..heads import FusedLateFusionHead
from ..utils.BaseModel import JointLogitsBaseModel
from ._common import MLP, build_hidden, build_siglip, hidden_features


class FusionNet(nn.Module):
    def __init__(self, num_classes, loss_fn, args=None):
        super().__init__()
        self.num_classes = num_classes
        self.loss_fn = loss_fn
        self.model = build_siglip(args)
        self.x1_model = MLP(input_dim=768, hidden_dim=512, num_classes=num_classes)
        self.x2_model = MLP(input_dim=768, hidden_dim=512, num_classes=num_classes)
        self.hidden = build_hidden(args)
        self.fused = FusedLateFusionHead(num_classes, mode="jlogits",
                                         precision=getattr(args, "head_precision", "auto"))

    def forward(self, x1_data, x2_data, label):
        output = self.model(x1_data, x2_data)
        h1, h2 = hidden_features(self, output['text_embeds'], output['image_embeds'])
        return self.fused(h1, h2, self.x1_model.classifier, self.x2_model.classifier, label)


class MultimodalFoodModel(JointLogitsBaseModel):
    def configure_optimizers(self):
        optimizer = self._sgd()
        if self.args.use_scheduler:
            scheduler = {'scheduler': StepLR(optimizer, step_size=50, gamma=0.5), 'interval': 'epoch', 'frequency': 1}
            return [optimizer], [scheduler]
        return optimizer

    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss(), args=self.args)
