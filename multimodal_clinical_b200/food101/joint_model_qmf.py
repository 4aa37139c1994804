"""Food101 QMF late fusion on SigLIP embeddings (food101/joint_model_qmf.py of the reference).  The fused
step starts at the last Linear(512, 101) of each modality's MLP; the first two layers and the dropout stay
in PyTorch (SURVEY.md §7 'hard parts')."""
import torch
import torch.nn as nn
from torch.optim.lr_scheduler import StepLR

from ..existing_algos.QMF import QMF
from ..heads import FusedLateFusionHead
from ..utils.BaseModel import QMFBaseModel
from ._common import MLP, build_hidden, build_siglip, hidden_features


class FusionNet(nn.Module):
    def __init__(self, args, loss_fn):
        super().__init__()
        self.args = args
        self.num_classes = self.args.num_classes
        self.num_modality = 2
        self.qmf = QMF(self.num_modality, self.args.num_samples)
        self.loss_fn = loss_fn
        self.model = build_siglip(args)
        self.x1_model = MLP(input_dim=768, hidden_dim=512, num_classes=self.num_classes)
        self.x2_model = MLP(input_dim=768, hidden_dim=512, num_classes=self.num_classes)
        self.hidden = build_hidden(args)
        self.fused = FusedLateFusionHead(self.num_classes, mode="qmf", n_data=self.args.num_samples,
                                         precision=getattr(args, "head_precision", "auto"))
        self.fused.bind_qmf(self.qmf)

    def forward(self, x1_data, x2_data, label, idx):
        output = self.model(x1_data, x2_data)
        h1, h2 = hidden_features(self, output['text_embeds'], output['image_embeds'])
        return self.fused(h1, h2, self.x1_model.classifier, self.x2_model.classifier, label, idx)


class MultimodalFoodModel(QMFBaseModel):
    def __init__(self, args):
        super().__init__(args)

    def configure_optimizers(self):
        optimizer = self._sgd()
        if self.args.use_scheduler:
            scheduler = {'scheduler': StepLR(optimizer, step_size=50, gamma=0.5), 'interval': 'epoch', 'frequency': 1}
            return [optimizer], [scheduler]
        return optimizer

    def _build_model(self):
        return FusionNet(args=self.args, loss_fn=nn.CrossEntropyLoss())
