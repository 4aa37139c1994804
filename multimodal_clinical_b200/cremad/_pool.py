"""Feature-side pooling in front of the fused heads (cremad/joint_model_qmf.py:48-55 of the reference): the global average
of the encoders' feature maps over space and, for the frame-stacked visual stream, over the T frames of a clip.  One
HBM-streaming CUDA kernel per map (csrc/lf_pool.cu) instead of view / permute / adaptive_avg_pool{2,3}d / flatten."""
import torch

from .. import _lib


class _PoolMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, batch):
        if not maps.is_cuda:
            raise _lib.LfError("pool_features runs on CUDA only (sm_100a); got a CPU tensor")
        if maps.dtype not in (torch.float32, torch.bfloat16):
            maps = maps.float()
        maps = maps.contiguous()
        n, C, H, W = maps.shape
        if n % batch:
            raise ValueError(f"{n} feature maps do not split into {batch} clips")
        T = n // batch
        out = torch.empty(batch, C, device=maps.device, dtype=maps.dtype)
        _lib.check(_lib.load().lf_pool_mean(maps.data_ptr(), out.data_ptr(), batch, T, C, H * W, maps.element_size(),
                                            torch.cuda.current_stream().cuda_stream), "lf_pool_mean")
        ctx.shape, ctx.batch = (n, C, H, W), batch
        return out

    @staticmethod
    def backward(ctx, g):
        n, C, H, W = ctx.shape
        g = g.contiguous()
        din = torch.empty(ctx.shape, device=g.device, dtype=g.dtype)
        _lib.check(_lib.load().lf_pool_mean_backward(g.data_ptr(), din.data_ptr(), ctx.batch, n // ctx.batch, C, H * W, g.element_size(),
                                                     torch.cuda.current_stream().cuda_stream), "lf_pool_mean_backward")
        return din, None


def pool_features(a, v):
    """Audio (B,512,H,W) and visual (B*T,512,H,W) feature maps -> (B,512) each: global average over space, and over
    the T frames (cremad/joint_model_qmf.py:48-55: view(B,-1,C,H,W).permute(0,2,1,3,4), adaptive_avg_pool2d/3d, flatten)."""
    B = a.size()[0]
    return _PoolMean.apply(a, B), _PoolMean.apply(v, B)
