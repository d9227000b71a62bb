"""`Trainer` with the reference's constructor arguments and `train()` loop shape (model/trainer.py:93-330, 521-690, 1081-1330),
driving `train.TrainEngine` instead of accelerate + autograd + torch.optim.

What is mirrored: batching ("sample" / "frame" with DynamicBatchSampler), per-process sharding of the batch list, warm-up + linear
decay (`num_warmup_updates * num_processes`, trainer.py:1179-1188), gradient accumulation, clip + AdamW(betas 0.9 / 0.98, eps 1e-8)
+ EMA on the main process, checkpoint naming / rotation / resume (`model_<update>.pt`, `model_last.pt`, `keep_last_n_checkpoints`,
`resumable_with_seed`), all in the reference's checkpoint format.
What is not (control plane, SURVEY.md §8 out of scope): wandb / tensorboard logging, sample generation during training, the
duration-predictor side loss and its alignment manager, bitsandbytes optimizers, `noise_scheduler`."""
from __future__ import annotations

import math
import os

import torch
from torch.utils.data import DataLoader, Dataset, SequentialSampler

from ..data import DynamicBatchSampler, collate_token_major, shard_batches
from ..optim import EmaSchedule, WarmupLinearDecay
from ..train import TrainEngine


def _exists(v):
    return v is not None


class Trainer:
    def __init__(self, model, epochs, learning_rate, weight_decay=0.1, num_warmup_updates=20000, save_per_updates=1000,
                 keep_last_n_checkpoints: int = -1, checkpoint_path=None, batch_size_per_gpu=32, batch_size_type: str = "sample",
                 max_samples=32, grad_accumulation_steps=1, max_grad_norm=1.0, noise_scheduler: str | None = None,
                 duration_predictor=None, logger: str | None = None, wandb_project="test_f5-tts", wandb_run_name="test_run",
                 wandb_resume_id: str = None, log_samples: bool = False, last_per_updates=None, accelerate_kwargs: dict = dict(),
                 ema_kwargs: dict = dict(), bnb_optimizer: bool = False, mel_spec_type: str = "vocos", is_local_vocoder: bool = False,
                 local_vocoder_path: str = "", model_cfg_dict: dict = dict(), dropout: float = 0.0, **ignored):
        if _exists(noise_scheduler):
            raise NotImplementedError("noise_scheduler is unused by the reference's CFM.forward and is not built")
        if _exists(duration_predictor):
            raise NotImplementedError("the duration-predictor side loss (trainer.py:330-520) is not built; "
                                      "model.DurationPredictor runs the eval forward only")
        if bnb_optimizer:
            raise NotImplementedError("bitsandbytes 8-bit AdamW is not built: the optimizer is the fused fp32 AdamW + EMA kernel")
        if log_samples:
            raise NotImplementedError("sample logging during training is control plane and is not built")
        import torch.distributed as dist
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.num_processes = dist.get_world_size() if self.distributed else 1
        self.is_main = self.rank == 0
        self.logger = None  # "wandb" / "tensorboard" accepted for signature compatibility; nothing is logged
        self.model = model
        self.epochs, self.learning_rate = epochs, learning_rate
        self.num_warmup_updates, self.save_per_updates = num_warmup_updates, save_per_updates
        self.keep_last_n_checkpoints = keep_last_n_checkpoints
        self.last_per_updates = last_per_updates if _exists(last_per_updates) else save_per_updates  # trainer.py default
        self.checkpoint_path = checkpoint_path if _exists(checkpoint_path) else "ckpts/test_f5-tts"
        self.batch_size_per_gpu, self.batch_size_type, self.max_samples = batch_size_per_gpu, batch_size_type, max_samples
        self.grad_accumulation_steps, self.max_grad_norm = grad_accumulation_steps, max_grad_norm
        ema = {k: v for k, v in ema_kwargs.items() if k in ("beta", "update_after_step", "update_every", "inv_gamma", "power", "min_value")}
        self.engine = TrainEngine(model, lr=learning_rate, betas=(0.9, 0.98), eps=1e-8, weight_decay=weight_decay,
                                  max_grad_norm=max_grad_norm, with_ema=self.is_main, ema_schedule=EmaSchedule(**ema), dropout=dropout)
        if self.distributed:
            self.engine.broadcast_params(0)
        self.scheduler = None
        self.losses: list[float] = []  # one mean loss per optimizer update (what the reference sends to its logger)

    # ------------------------------------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, update, last=False):
        """trainer.py:521-598 (main process writes; numbered checkpoints rotate under keep_last_n_checkpoints)"""
        if self.distributed:
            torch.distributed.barrier()
        if not self.is_main:
            return
        os.makedirs(self.checkpoint_path, exist_ok=True)
        sched = dict(warmup=self.scheduler.warmup, decay=self.scheduler.decay, base_lr=self.scheduler.base_lr) if self.scheduler else None
        if last:
            self.engine.save_checkpoint(os.path.join(self.checkpoint_path, "model_last.pt"), update, sched)
            return
        if self.keep_last_n_checkpoints == 0:
            return
        self.engine.save_checkpoint(os.path.join(self.checkpoint_path, f"model_{update}.pt"), update, sched)
        if self.keep_last_n_checkpoints > 0:
            numbered = sorted((f for f in os.listdir(self.checkpoint_path)
                               if f.startswith("model_") and f.endswith(".pt") and f != "model_last.pt" and f[6:-3].isdigit()),
                              key=lambda f: int(f[6:-3]))
            while len(numbered) > self.keep_last_n_checkpoints:
                os.remove(os.path.join(self.checkpoint_path, numbered.pop(0)))

    def find_checkpoint(self):
        """trainer.py:600-646: model_last.pt first; else the training checkpoint model_<update>.{pt,safetensors} with the highest
        update; else the first (sorted) pretrained_*.{pt,safetensors} — the fine-tune entry point.  Returns a path or None."""
        import re
        if not _exists(self.checkpoint_path) or not os.path.isdir(self.checkpoint_path):
            return None
        if os.path.exists(os.path.join(self.checkpoint_path, "model_last.pt")):
            return os.path.join(self.checkpoint_path, "model_last.pt")
        files = [f for f in os.listdir(self.checkpoint_path)
                 if (f.startswith("model_") or f.startswith("pretrained_")) and f.endswith((".pt", ".safetensors"))]
        training = [f for f in files if f.startswith("model_") and f != "model_last.pt"]
        pretrained = [f for f in files if f.startswith("pretrained_")]
        name = None
        if training:
            best = -1
            for f in training:
                m = re.search(r"model_(\d+)", f)
                if m and int(m.group(1)) > best:
                    best, name = int(m.group(1)), f
        elif pretrained:
            name = sorted(pretrained)[0]
        if name is None:
            others = [f for f in os.listdir(self.checkpoint_path) if f.endswith((".pt", ".safetensors"))]
            if others:  # never start from random init silently next to weight files this search order does not pick up
                raise FileNotFoundError(f"{self.checkpoint_path} holds {others} but none is model_last.pt, model_<update>.* or pretrained_*")
            return None
        return os.path.join(self.checkpoint_path, name)

    def load_checkpoint(self) -> int:
        """trainer.py:600-827; returns the update to resume at (0 = fresh optimizer on loaded or random weights)"""
        path = self.find_checkpoint()
        if path is None:
            return 0
        return int(self.engine.load_checkpoint(path, grad_accumulation_steps=self.grad_accumulation_steps))

    # ------------------------------------------------------------------------------------------------ loop
    def _batches(self, train_dataset, resumable_with_seed):
        if self.batch_size_type == "sample":
            g = torch.Generator()
            if _exists(resumable_with_seed):
                g.manual_seed(resumable_with_seed)
            return None, g
        if self.batch_size_type == "frame":
            return DynamicBatchSampler(SequentialSampler(train_dataset), self.batch_size_per_gpu, max_samples=self.max_samples,
                                       random_seed=resumable_with_seed, drop_residual=False), None
        raise ValueError(f"batch_size_type must be either 'sample' or 'frame', but received {self.batch_size_type}")

    def _epoch_batches(self, train_dataset, sampler, gen, epoch):
        """this process's list of index batches for one epoch"""
        if sampler is not None:
            sampler.set_epoch(epoch)
            order = list(iter(sampler))
        else:
            perm = torch.randperm(len(train_dataset), generator=gen).tolist()
            bs = self.batch_size_per_gpu
            order = [perm[i:i + bs] for i in range(0, len(perm), bs)]
        return shard_batches(order, self.rank, self.num_processes)

    def train(self, train_dataset: Dataset, num_workers=16, resumable_with_seed: int = None):
        eng = self.engine
        sampler, gen = self._batches(train_dataset, resumable_with_seed)
        per_epoch = len(self._epoch_batches(train_dataset, sampler, torch.Generator().manual_seed(0) if gen is not None else None, 0))
        if per_epoch == 0:
            raise ValueError("the dataset yields no batch for this process")
        accum = self.grad_accumulation_steps
        warmup_updates = self.num_warmup_updates * self.num_processes  # trainer.py:1179-1181
        total_updates = math.ceil(per_epoch / accum) * self.epochs
        # accelerate steps the prepared scheduler num_processes times per optimizer update (see WarmupLinearDecay)
        self.scheduler = WarmupLinearDecay(self.learning_rate, warmup_updates, total_updates, steps_per_update=self.num_processes)
        start_update = self.load_checkpoint()
        global_update = start_update
        # the reference restores the scheduler's own state (one step per applied update) while it resumes the update COUNT at
        # `update + 1` (trainer.py:812): keep the two apart so the learning rate continues exactly where it stopped
        lr_update = max(start_update - 1, 0)
        skipped_epoch, skipped_batch = 0, 0
        if _exists(resumable_with_seed):
            start_step = start_update * accum
            skipped_epoch, skipped_batch = int(start_step // per_epoch), start_step % per_epoch
        import torch.distributed as dist
        if gen is not None:
            for _ in range(skipped_epoch):  # "sample" mode: replay the shuffles of the epochs already trained
                torch.randperm(len(train_dataset), generator=gen)
        for epoch in range(skipped_epoch, self.epochs):
            batches = self._epoch_batches(train_dataset, sampler, gen, epoch)
            if epoch == skipped_epoch and skipped_batch:
                batches = batches[skipped_batch:]
            loader = DataLoader(train_dataset, collate_fn=collate_token_major, batch_sampler=batches, num_workers=num_workers,
                                pin_memory=num_workers > 0, persistent_workers=False, **(dict(prefetch_factor=2) if num_workers > 0 else {}))
            micro, acc_loss = 0, None
            eng.zero_grad()

            def apply_update():
                nonlocal global_update, acc_loss, lr_update
                scale = eng.allreduce_grads() if self.distributed else 1.0
                eng.step(lr=self.scheduler.lr(lr_update), grad_scale=scale / accum)  # the loss is divided by accum either way
                eng.zero_grad()
                global_update += 1
                lr_update += 1
                self.losses.append(float(acc_loss) / accum)
                acc_loss = None
                if global_update % self.save_per_updates == 0:
                    self.save_checkpoint(global_update)
                if global_update % self.last_per_updates == 0:
                    self.save_checkpoint(global_update, last=True)

            for batch in loader:
                last = micro % accum == accum - 1
                loss, _, _ = eng.loss_and_grads(batch["mel"], batch["text"], lens=batch["mel_lengths"],
                                                overlap_allreduce=last and self.distributed)
                acc_loss = loss.clone() if acc_loss is None else acc_loss + loss
                micro += 1
                if last:
                    apply_update()
            if acc_loss is not None:  # accelerate's accumulate() also synchronises on the last batch of the dataloader
                apply_update()
        self.save_checkpoint(global_update, last=True)
        if self.distributed:
            dist.barrier()
        return global_update
