"""Synthetic image-text multitask batches with the schema of the reference collate functions
(vqa_clip_data.py:365-390 and the gqa / nlvr / caption twins) and the per-task batch ratios of
multitask.py:682-695 (vqa b, gqa int(b*100/60), nlvr int(b*20/60), caption int(b*50/60)).  No datasets or
tokenizers exist offline: token ids are uniform in [3, 50000), CLIP grid features N(0,1) on the 7x7
grid, boxes all-zero as the CLIP-feature loaders produce them (vqa_clip_data.py:198)."""
from __future__ import annotations

from typing import Dict, List

import torch

# task -> (text length, answer/target length, images per sample); text 20 tokens (--max_text_length 20), caption
# prompts are short and its targets up to 40 tokens (multitask.py:682-695)
TASK_SHAPES = {"vqa": (20, 5, 1), "gqa": (20, 5, 1), "nlvr": (20, 3, 2), "caption": (4, 40, 1),
               # video-text multitask (multitask_video.py:738-743: 64 CLIP-ViT frame features of width 512, text up to 600
               # tokens -- fixed at 256 here so that text + frames = 320 tokens, targets <= 20; equal batch per task)
               "tvqa": (256, 20, 1), "how2qa": (256, 20, 1), "tvc": (256, 20, 1), "yc2c": (256, 20, 1)}
_TASK_INDEX = {"vqa": 0, "gqa": 1, "nlvr": 2, "caption": 3, "tvqa": 4, "how2qa": 5, "tvc": 6, "yc2c": 7}
VIDEO_TASKS = ["tvqa", "how2qa", "tvc", "yc2c"]


def task_batch_sizes(batch_size: int, tasks=None) -> Dict[str, int]:
    if tasks is not None and all(t in VIDEO_TASKS for t in tasks):
        return {t: batch_size for t in tasks}               # multitask_video.py:740-743: no ratio scaling
    return {"vqa": batch_size, "gqa": int(batch_size * 100 / 60), "nlvr": int(batch_size * 20 / 60),
            "caption": int(batch_size * 50 / 60)}


def make_task_batch(task: str, B: int, feat_dim: int = 2048, grid: int = 49, seed: int = 0, vocab_hi: int = 50000,
                    pin: bool = False) -> Dict:
    """One HOST batch (CPU tensors, optionally pinned) for ``task``."""
    g = torch.Generator().manual_seed(seed * 1000003 + _TASK_INDEX[task])
    Lt, T, n_img = TASK_SHAPES[task]
    ids = torch.randint(3, vocab_hi, (B, Lt), generator=g, dtype=torch.int64)
    tgt = torch.randint(3, vocab_hi, (B, T), generator=g, dtype=torch.int64)
    if T > 2:                                   # ragged targets: -100 padding after a per-sample length >= 2
        lens = torch.randint(2, T + 1, (B, 1), generator=g)
        tgt = torch.where(torch.arange(T).view(1, T) < lens, tgt, torch.full_like(tgt, -100))
    if n_img == 2:
        feats = torch.randn(B, 2, grid, feat_dim, generator=g)
        boxes = torch.zeros(B, 2, grid, 4)
    else:
        feats = torch.randn(B, grid, feat_dim, generator=g)
        boxes = torch.zeros(B, grid, 4)
    batch = {"task": task, "input_ids": ids, "vis_feats": feats, "boxes": boxes, "target_ids": tgt}
    if task in ("vqa", "gqa", "tvqa", "how2qa"):
        batch["scores"] = torch.ones(B)
    if pin:
        batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
    return batch


def multitask_cycle(batch_size: int, tasks: List[str], feat_dim: int = 2048, seed: int = 0, pin: bool = False,
                    rank: int = 0, world: int = 1, vocab_hi: int = 50000, grid: int = 49) -> List[Dict]:
    """One round-robin cycle over ``tasks`` at the reference ratios; with world > 1 each batch is this rank's
    contiguous shard of the GLOBAL task batch (strong scaling: the global batch is fixed)."""
    sizes = task_batch_sizes(batch_size, tasks)
    out = []
    for t in tasks:
        gb = make_task_batch(t, sizes[t], feat_dim=feat_dim, seed=seed, vocab_hi=vocab_hi, grid=grid)
        b = shard_batch(gb, rank, world)
        if pin:
            b = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()}
        out.append(b)
    return out


def shard_batch(batch: Dict, rank: int, world: int) -> Dict:
    """Contiguous split by sample (what DistributedSampler intends, vqa_clip_data.py:409-410).  Every rank gets
    ceil(B/world) or floor(B/world) samples; shapes may differ by one sample between ranks."""
    if world == 1:
        return batch
    B = batch["input_ids"].shape[0]
    lo = (B * rank) // world
    hi = (B * (rank + 1)) // world
    return {k: (v[lo:hi].contiguous() if torch.is_tensor(v) else v) for k, v in batch.items()}


def batch_nbytes(batch: Dict) -> int:
    return sum(v.numel() * v.element_size() for v in batch.values() if torch.is_tensor(v))
