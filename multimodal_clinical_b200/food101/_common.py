"""Shared pieces of the Food101 models: the per-modality MLP whose last Linear is the fused head's
classifier, and the SigLIP feature producer (food101/joint_model_qmf.py:12-26, 42-64 of the reference)."""
import torch.nn as nn


class MLP(nn.Module):
    """768 -> 512 -> 512 -> num_classes with ReLU + Dropout(0.2); parameter names ``mlp.{0,3,6}`` as in the
    reference.  ``hidden()`` runs everything but the last Linear (``mlp.6``), which the fused step owns."""

    def __init__(self, input_dim=768, hidden_dim=512, num_classes=101):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.2),
            nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.2),
            nn.Linear(hidden_dim, num_classes))

    def hidden(self, x):
        return self.mlp[:6](x)

    @property
    def classifier(self):
        return self.mlp[6]

    def forward(self, x):
        return self.mlp(x)


class PrecomputedEmbeddings(nn.Module):
    """Feature producer used when SigLIP weights are not available offline: the batch already holds the
    768-d text / image embeddings and they are passed through."""

    def forward(self, x1, x2):
        return {"text_embeds": x1, "image_embeds": x2}


def build_siglip(args=None):
    name = getattr(args, "encoder", "google/siglip-base-patch16-224") if args is not None else "google/siglip-base-patch16-224"
    if name in ("precomputed", "synthetic", None):
        return PrecomputedEmbeddings()
    from transformers import AutoModel
    model = AutoModel.from_pretrained(name)
    for param in model.parameters():
        param.requires_grad = True
    return model
