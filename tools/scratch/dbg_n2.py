"""torchrun --nproc-per-node 2 tools/scratch/dbg_n2.py : sharded vs single mean-fusion step, header statistics of both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from multimodal_clinical_b200.step import LateFusionStep
from multimodal_clinical_b200._lib import STAT
C, D, B = 309, 512, 2048
Bg = B * world
for overlap in (False, True):
    for sgd in (False, True):
        eng_s = LateFusionStep(C, mode="jlogits", device=dev, precision="bf16")
        eng_1 = LateFusionStep(C, mode="jlogits", device=dev, precision="bf16", sharded=False)
        eng_s.cal_overlap = overlap; eng_1.cal_overlap = overlap
        g = torch.Generator().manual_seed(5)
        Ws = [(torch.rand(C, D, generator=g) * 0.08 - 0.04).to(dev) for _ in range(2)]; bs = [(torch.rand(C, generator=g) * 0.08 - 0.04).to(dev) for _ in range(2)]
        W1 = [x.clone() for x in Ws]; b1 = [x.clone() for x in bs]
        if sgd:
            eng_s.enable_sgd(lr=1e-2); eng_1.enable_sgd(lr=1e-2)
        for s in range(3):
            gg = torch.Generator(device=dev).manual_seed(100 + s)
            f1 = torch.randn(Bg, D, generator=gg, device=dev).bfloat16(); f2 = torch.randn(Bg, D, generator=gg, device=dev).bfloat16()
            y = torch.randint(0, C, (Bg,), generator=gg, device=dev)
            sl = slice(rank * B, (rank + 1) * B)
            o1 = eng_1.step([f1, f2], W1, b1, y, ogm_alpha=0.8)
            torch.cuda.synchronize()
            os_ = eng_s.step([f1[sl], f2[sl]], Ws, bs, y[sl], ogm_alpha=0.8)
            torch.cuda.synchronize()
            h1 = o1.stats[:12].cpu().tolist(); hs = os_.stats[:12].cpu().tolist()
            print(f"[r{rank}] overlap={overlap} sgd={sgd} step={s} loss1={float(o1.loss):.6g} losss={float(os_.loss):.6g}\n   single {['%.6g' % v for v in h1]}\n   shard  {['%.6g' % v for v in hs]}\n   ema1 {eng_1.ema_x[0,:3].tolist()} emas {eng_s.ema_x[0,:3].tolist()} coeff {eng_1.coeff.tolist()} {eng_s.coeff.tolist()}", flush=True)
        eng_s.close()
if world > 1:
    dist.destroy_process_group()
