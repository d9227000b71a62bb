"""N-rank check of model.Trainer (run under torchrun on N GPUs): every rank trains on its shard of the batch list with the bucketed
all-reduce launched under the backward; afterwards the parameters of all ranks must be bit-identical, every rank must have made the
same number of updates, and only rank 0 writes checkpoints."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_cfm  # noqa: E402
from oracle import f5_oracle as O  # noqa: E402
from test_gpu_train import _ToyDataset  # noqa: E402
from eraxvif5tts_b200.model import Trainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = O.DiTConfig.tiny()
model, _ = build_cfm(cfg, seed=rank, device=f"cuda:{local}")  # different initial weights per rank: Trainer's broadcast must fix that
model.vocab_char_map = {chr(97 + i): i for i in range(26)}
ckdir = os.path.join(tempfile.gettempdir(), "f5b_trainer_ddp_check")
if rank == 0:
    import shutil
    shutil.rmtree(ckdir, ignore_errors=True)
dist.barrier()
tr = Trainer(model, epochs=2, learning_rate=2e-3, weight_decay=0.0, num_warmup_updates=1, save_per_updates=2, keep_last_n_checkpoints=2,
             checkpoint_path=ckdir, batch_size_per_gpu=300, batch_size_type="frame", max_samples=6, grad_accumulation_steps=2)
updates = tr.train(_ToyDataset(n_items=40), num_workers=0, resumable_with_seed=11)
flat = tr.engine.p
digest = torch.stack([flat.double().sum(), flat.double().abs().sum(), torch.tensor(float(updates), device=flat.device, dtype=torch.float64)])
all_d = [torch.empty_like(digest) for _ in range(world)]
dist.all_gather(all_d, digest)
same = all(torch.equal(all_d[0], d) for d in all_d)
ref = flat.clone()
dist.broadcast(ref, 0)
bit_identical = bool(torch.equal(ref, flat))
ok = torch.tensor([int(same and bit_identical)], device=flat.device)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    files = sorted(os.listdir(ckdir))
    print(f"world {world}: updates {updates}, losses {tr.losses[0]:.4f} -> {tr.losses[-1]:.4f}, params bit-identical on all ranks: {bool(ok.item())}, "
          f"checkpoints {files}")
    assert bool(ok.item()) and "model_last.pt" in files and tr.losses[-1] < tr.losses[0]
dist.destroy_process_group()
