"""2+ rank check of the data-parallel training step (run under torchrun on N GPUs):
  * every rank computes gradients on ITS shard, one flat NCCL all-reduce averages them;
  * rank 0 also accumulates the gradients of ALL shards locally (gradient accumulation) -> the two must agree;
  * after one fused AdamW step from identical (broadcast) weights, the parameters of all ranks are bit-identical."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_cfm  # noqa: E402
from oracle import f5_oracle as O  # noqa: E402
from eraxvif5tts_b200.train import TrainEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = O.DiTConfig.tiny()
model, _ = build_cfm(cfg, seed=rank, device=f"cuda:{local}")  # different initial weights per rank: the broadcast must fix that
eng = TrainEngine(model, lr=1e-3)
eng.broadcast_params(0)
B, n = 2, 128


def shard(r):
    g = torch.Generator().manual_seed(100 + r)
    x1 = (torch.randn(B, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5)
    x0 = torch.randn(B, n, cfg.mel_dim, generator=g)
    time = torch.rand(B, generator=g)
    text = torch.randint(0, cfg.text_num_embeds, (B, 20), generator=g)
    span = torch.zeros(B, n, dtype=torch.bool)
    span[:, 30:100] = True
    return x1, text, dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False)


x1, text, dr = shard(rank)
eng.zero_grad()
eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dr, overlap_allreduce=os.environ.get("F5B_ALLREDUCE_OVERLAP", "1") != "0", buckets=2)
scale = eng.allreduce_grads()
avg = eng.g * scale
ok = True
if rank == 0:
    ref_eng_g = torch.zeros_like(eng.g)
    for r in range(world):
        eng.zero_grad()
        x1r, textr, drr = shard(r)
        eng.loss_and_grads(x1r.cuda(), textr.cuda(), draws=drr)
        eng._fold_split_grads()
        ref_eng_g += eng.g
    ref = ref_eng_g / world
    err = float((avg - ref).norm() / ref.norm())
    print(f"[ddp] all-reduced gradient vs local accumulation over {world} shards: rel err {err:.3e}")
    ok = err < 2e-3
    eng.g.copy_(avg / scale)
dist.barrier()
eng.step(grad_scale=scale)
torch.cuda.synchronize()
digest = torch.stack([eng.p.double().sum(), eng.p.double().abs().sum()])
all_d = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(all_d, digest)
same = all(torch.equal(all_d[0], d) for d in all_d)
if rank == 0:
    print(f"[ddp] parameters identical on all {world} ranks after the step: {same}")
    print("[ddp] OK" if (ok and same) else "[ddp] FAILED")
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
