"""Shared pieces of the Food101 models: the per-modality MLP whose last Linear is the fused head's
classifier, and the SigLIP feature producer (food101/joint_model_qmf.py:12-26, 42-64 of the reference)."""
import torch.nn as nn

from ..hidden import FusedHiddenPair


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


class FusedMLPHidden(nn.Module):
    """Both modalities' ``mlp[0:6]`` (Linear-ReLU-Dropout twice, food101/joint_model_qmf.py:15-20) on the fused hidden-layer
    kernels: two launches of the tensor-pipe GEMM with bias + ReLU + dropout in the epilogue instead of 2 x (cuBLAS GEMM +
    three elementwise kernels) per modality.  Parameter-free; dropout probability and train / eval mode are read from the
    MLPs' own ``nn.Dropout`` modules."""

    def __init__(self, precision: str = "auto"):
        super().__init__()
        self.layer1 = FusedHiddenPair(precision=precision)
        self.layer2 = FusedHiddenPair(precision=precision)

    def forward(self, mlp1: "MLP", mlp2: "MLP", e1, e2):
        for layer, k in ((self.layer1, 2), (self.layer2, 5)):
            layer.drop_p = float(mlp1.mlp[k].p)
            layer.train(mlp1.mlp[k].training)
        h1, h2 = self.layer1(e1, e2, mlp1.mlp[0], mlp2.mlp[0])
        return self.layer2(h1, h2, mlp1.mlp[3], mlp2.mlp[3])


def hidden_features(net, e1, e2):
    """(h1, h2) in front of the fused head: the fused hidden layers, or -- ``args.fused_hidden: false`` -- the MLPs' own
    PyTorch layers (cuBLAS), which is what round 1 shipped."""
    if getattr(net, "hidden", None) is not None:
        return net.hidden(net.x1_model, net.x2_model, e1, e2)
    return net.x1_model.hidden(e1), net.x2_model.hidden(e2)


def build_hidden(args):
    return FusedMLPHidden(precision=getattr(args, "head_precision", "auto")) if getattr(args, "fused_hidden", True) else None


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
