"""Wire format between the reference's dataset classes and ``train_step`` / ``test_step``.

``collate`` builds, from per-sample entries shaped like the reference's ``__getitem__`` results, exactly the batch dict the
reference's per-task ``collate_fn`` builds (vqa_clip_data.py:300-390, gqa_clip_data.py:231-324, nlvr_clip_data.py:182-255,
caption_clip_data.py:275-352, video/{tvqa,how2qa,tvc,yc2c}_data.py ``collate_fn``): right-padded ``input_ids``, ``target_ids``
whose padding is -100, fp32 ``vis_feats`` / ``boxes`` (two images per sample for NLVR), and each task's bookkeeping lists under
the reference's key names.  ``resize_frames`` is the video loaders' ``resize`` to ``n_boxes`` frames
(video/tvqa_data.py:33-46) and ``MultitaskLoader`` the per-epoch task schedule of multitask_data.py:5-64.  Pinned by
``tests/golden/collate_cases.json`` (the unmodified reference functions run on seeded entries).

Host plumbing only: CPU tensors in, CPU tensors out (``pin=True`` for the trainer's asynchronous copies); the dataset
readers themselves (h5 / npz feature files, tokenizers) need files this offline image does not have and are not restated.
"""
from __future__ import annotations

import random
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn.functional as F

IMAGE_TASKS = ("vqa", "gqa", "nlvr", "caption")
VIDEO_TASKS = ("tvqa", "how2qa", "tvc", "yc2c")

# bookkeeping lists per task: batch key -> entry key.  A list only collects the entries that carry the key (the
# reference appends under ``if key in entry``), so inference entries without answers give empty lists.
_LISTS = {
    "vqa": {"sent": "sent", "question_ids": "question_id", "answers": "answer", "all_answers": "all_answers", "labels": "label"},
    "gqa": {"sent": "sent", "question_ids": "question_id", "answers": "answer", "all_answers": "all_answers",
            "all_answers_tokenized": "all_answers_tokenized", "best_answers_tokenized": "best_answers_tokenized", "labels": "label"},
    "nlvr": {"sent": "sent", "question_ids": "question_id", "answers": "answer"},
    "caption": {"img_id": "img_id", "input_text": "input_text", "targets": "targets"},
    "tvqa": {"sent": "sent", "question_ids": "question_id", "answers": "answer"},
    "how2qa": {"sent": "sent", "question_ids": "question_id", "answers": "answer"},
    "tvc": {"sent": "sent", "question_ids": "question_id", "answers": "answer", "video_ids": "video_id", "tss": "ts"},
    "yc2c": {"sent": "sent", "question_ids": "question_id", "answers": "answer", "video_ids": "video_id"},
}


def _ids(x) -> torch.Tensor:
    return torch.as_tensor(x, dtype=torch.long).reshape(-1)


def _pad_ids(rows: Sequence[torch.Tensor], fill: int) -> torch.Tensor:
    width = max(int(r.numel()) for r in rows)
    out = torch.full((len(rows), width), fill, dtype=torch.long)
    for i, r in enumerate(rows):
        out[i, :r.numel()] = r
    return out


def collate(task: str, entries: Sequence[Dict], pad_token_id: int, pin: bool = False) -> Dict:
    """Per-sample ``entries`` -> the reference's batch dict for ``task``.

    An entry carries ``input_ids`` (any int sequence), ``vis_feats`` ``[V_L, feat_dim]`` and ``boxes`` ``[V_L, 4]``
    (``[2, V_L, ...]`` for NLVR; tensors, arrays or nested lists), optionally ``target_ids``, and the task's text fields.
    All samples of a batch share ``V_L``; caption batches also get the all-ones ``vis_attention_mask`` over each sample's
    ``n_boxes`` leading rows (caption_clip_data.py:289-316)."""
    if task not in _LISTS:
        raise ValueError(f"unknown task {task!r}: expected one of {sorted(_LISTS)}")
    if len(entries) == 0:
        raise ValueError("collate needs at least one entry")
    first = entries[0]
    batch: Dict = {"input_ids": _pad_ids([_ids(e["input_ids"]) for e in entries], pad_token_id)}
    if "target_ids" in first:
        tgt = _pad_ids([_ids(e["target_ids"]) for e in entries], pad_token_id)
        # the reference masks by value, not by length: a pad id inside a target is ignored by the loss as well
        batch["target_ids"] = tgt.masked_fill(tgt == pad_token_id, -100)
    feats = [torch.as_tensor(e["vis_feats"], dtype=torch.float32) for e in entries]
    boxes = [torch.as_tensor(e["boxes"], dtype=torch.float32) for e in entries]
    want = 3 if task == "nlvr" else 2
    for f, b in zip(feats, boxes):
        if f.dim() != want or b.shape != f.shape[:-1] + (4,) or f.shape != feats[0].shape:
            raise ValueError(f"{task}: vis_feats {tuple(f.shape)} / boxes {tuple(b.shape)} do not form one "
                             f"[B{', 2' if task == 'nlvr' else ''}, V_L, feat_dim] batch")
    batch["boxes"] = torch.stack(boxes)
    batch["vis_feats"] = torch.stack(feats)
    if task == "caption":
        V_L = feats[0].shape[0]
        n = torch.tensor([int(e.get("n_boxes", V_L)) for e in entries]).view(-1, 1)
        batch["vis_attention_mask"] = (torch.arange(V_L).view(1, -1) < n).float()
        batch["img_paths"] = []
    for key, src in _LISTS[task].items():
        batch[key] = [e[src] for e in entries if src in e]
    if task in ("vqa", "gqa"):
        batch["scores"] = torch.tensor([float(e["score"]) for e in entries if "score" in e], dtype=torch.float32)
    if task == "nlvr" and first.get("label") is not None:
        batch["labels"] = torch.tensor([int(e["label"]) for e in entries], dtype=torch.long)
    batch["task"] = task
    if pin:
        batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
    return batch


def resize_frames(feats: torch.Tensor, length: int) -> torch.Tensor:
    """``[L, D]`` frame features -> ``[length, D]``: zero rows appended when the clip is short, an adaptive max over time
    windows when it is long (video/tvqa_data.py:33-46; ``n_boxes`` = 64 in the video scripts)."""
    L, D = feats.shape
    if L < length:
        return torch.cat([feats, feats.new_zeros(length - L, D)], dim=0)
    if L == length:
        return feats
    return F.adaptive_max_pool1d(feats.t().unsqueeze(0), length).squeeze(0).t()


class MultitaskLoader:
    """One iterator over several task loaders (multitask_data.py:5-64).  Each loader has ``.task``, ``len()`` (batches per
    epoch), ``iter()`` and, optionally, ``.sampler.set_epoch``.  ``roundrobin``: every batch of every loader once per epoch;
    ``balanced``: ``n_batches`` (default: the mean loader length) per task.  The epoch's task list is shuffled by
    ``random.Random(epoch)`` -- the same order on every rank without communication -- and consumed from its END, as the
    reference pops it."""

    def __init__(self, loaders: Iterable, shuffle: bool = True, sampling: str = "roundrobin", n_batches: Optional[int] = None):
        self.loaders = list(loaders)
        if sampling not in ("roundrobin", "balanced"):
            raise ValueError(f"sampling must be 'roundrobin' or 'balanced', got {sampling!r}")
        self.task2loader = {ld.task: ld for ld in self.loaders}
        self.task2len = {ld.task: len(ld) for ld in self.loaders}
        self.shuffle, self.sampling, self.n_batches = shuffle, sampling, n_batches
        self.epoch_tasks: List[str] = []
        self.set_epoch(0)

    def set_epoch(self, epoch: int) -> None:
        for ld in self.loaders:
            sampler = getattr(ld, "sampler", None)
            if sampler is not None and hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)
        if self.sampling == "roundrobin":
            counts = dict(self.task2len)
        else:
            n = self.n_batches if self.n_batches is not None else sum(self.task2len.values()) // len(self.loaders)
            counts = {t: n for t in self.task2len}
        tasks = [t for t, n in counts.items() for _ in range(n)]
        if self.shuffle:
            random.Random(epoch).shuffle(tasks)
        self.epoch_tasks = tasks

    def __iter__(self):
        self._iters = {t: iter(ld) for t, ld in self.task2loader.items()}
        return self

    def __next__(self):
        if not self.epoch_tasks:
            raise StopIteration
        return next(self._iters[self.epoch_tasks.pop()])

    def __len__(self) -> int:
        return len(self.epoch_tasks)


class TaskBatches:
    """The smallest loader ``MultitaskLoader`` accepts: fixed-size batches of one task's entries through ``collate``, this
    rank's samples only (every ``world``-th sample starting at ``rank``, the split ``DistributedSampler`` makes without
    shuffling; with ``shuffle`` the order is re-drawn from ``seed + epoch`` as ``DistributedSampler.set_epoch`` does)."""

    def __init__(self, task: str, entries: Sequence[Dict], batch_size: int, pad_token_id: int, rank: int = 0, world: int = 1,
                 shuffle: bool = False, seed: int = 0, drop_last: bool = False, pin: bool = False):
        self.task, self.entries, self.batch_size, self.pad_token_id = task, entries, batch_size, pad_token_id
        self.rank, self.world, self.shuffle, self.seed, self.drop_last, self.pin = rank, world, shuffle, seed, drop_last, pin
        self.sampler = self
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def _indices(self) -> List[int]:
        n = len(self.entries)
        order = list(range(n))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, generator=g).tolist()
        per_rank = -(-n // self.world)
        order += order[:per_rank * self.world - n]              # wrap-around padding so every rank gets per_rank samples
        return order[self.rank::self.world]

    def __len__(self) -> int:
        n = -(-len(self.entries) // self.world)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        idx = self._indices()
        for b in range(len(self)):
            chunk = idx[b * self.batch_size:(b + 1) * self.batch_size]
            yield collate(self.task, [self.entries[i] for i in chunk], self.pad_token_id, pin=self.pin)
