import torch
import torch.nn.functional as F


def pool_features(a, v):
    """Audio (B,512,H,W) and visual (B*T,512,H,W) feature maps -> (B,512) each
    (cremad/joint_model_qmf.py:48-55: global average over space, and over the T frames)."""
    (_, C, H, W) = v.size()
    B = a.size()[0]
    v = v.view(B, -1, C, H, W).permute(0, 2, 1, 3, 4)
    a = torch.flatten(F.adaptive_avg_pool2d(a, 1), 1)
    v = torch.flatten(F.adaptive_avg_pool3d(v, 1), 1)
    return a, v
