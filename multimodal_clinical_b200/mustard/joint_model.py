"""MUStARD sarcasm detection, three modalities (mustard/joint_model.py of the reference): an LSTM encoder per modality,
mean fusion of the three two-way heads.  The heads (``x{1,2,3}_model.fc3``, same state-dict names) and the loss run on the
fused multi-head kernel; the encoders up to ``relu(fc2(.))`` stay in PyTorch as feature producers."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..multi import FusedMeanFusionHeads
from ..utils.MultiModel import MeanFusionMultiBaseModel


class LstmClassifier(nn.Module):
    hidden_dim = 384

    def __init__(self, input_dim, num_classes):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, self.hidden_dim)
        self.lstm = nn.LSTM(self.hidden_dim, self.hidden_dim, batch_first=True)
        self.fc2 = nn.Linear(self.hidden_dim, 100)
        self.fc3 = nn.Linear(100, num_classes)

    def embed(self, x):
        """(B, S, input_dim) -> (B, 100): what reaches fc3 in the reference (mustard/joint_model.py:24-41)."""
        B, S, H = x.shape
        x = self.fc1(x.reshape(-1, H)).view(B, S, -1)
        _, (hn, _) = self.lstm(x)
        return F.relu(self.fc2(hn[-1]))

    def forward(self, x):
        return self.fc3(self.embed(x))


class FusionNet(nn.Module):
    def __init__(self, num_classes, loss_fn):
        super().__init__()
        self.x1_model = LstmClassifier(371, num_classes)
        self.x2_model = LstmClassifier(81, num_classes)
        self.x3_model = LstmClassifier(300, num_classes)
        self.num_classes = num_classes
        self.loss_fn = loss_fn
        self.fused = FusedMeanFusionHeads(num_classes)

    def forward(self, x1_data, x2_data, x3_data, label):
        """-> (x1_logits, x2_logits, x3_logits, avg_logits, loss)   mustard/joint_model.py:72-83"""
        models = (self.x1_model, self.x2_model, self.x3_model)
        feats = [m.embed(x) for m, x in zip(models, (x1_data, x2_data, x3_data))]
        return self.fused(feats, [m.fc3 for m in models], label.flatten())


class MultimodalMustardModel(MeanFusionMultiBaseModel):
    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss())
