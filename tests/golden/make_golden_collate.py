"""Golden fixture for SURVEY section 8 row f-4 (data / wire formats): what the reference's dataset classes hand to
``train_step`` / ``test_step``.  Test infrastructure; run in the build container:

    python tests/golden/make_golden_collate.py

The unmodified ``collate_fn`` of every task the shipped scripts train on (vqa_clip_data.py:300-390, gqa_clip_data.py:231-324,
nlvr_clip_data.py:182-255, caption_clip_data.py:275-352, video/{tvqa,how2qa,tvc,yc2c}_data.py), the video loaders' ``resize``
to ``n_boxes`` frames (video/tvqa_data.py:33-46) and ``MultitaskLoader``'s per-epoch task order (multitask_data.py:5-64) are
run on small seeded entries; entries and results go to ``collate_cases.json``.  Packages the data modules import but the
collate path never touches (h5py, more_itertools, language_evaluation, sacrebleu, ftfy, timm) are absent offline and
replaced by empty modules.
"""
import json
import os
import random
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import as R  # noqa: E402

PAD = 1
IMAGE_TASKS = ["vqa", "gqa", "nlvr", "caption"]
VIDEO_TASKS = ["tvqa", "how2qa", "tvc", "yc2c"]


def _import_data_modules():
    R.install_shims()
    for name in ("h5py", "more_itertools", "language_evaluation", "sacrebleu", "ftfy", "timm"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, os.path.join(R.REF_SRC, "video"))
    import vqa_clip_data, gqa_clip_data, nlvr_clip_data, caption_clip_data, multitask_data  # noqa: E401
    import tvqa_data, how2qa_data, tvc_data, yc2c_data  # noqa: E401
    classes = {"vqa": vqa_clip_data.VQAFineTuneDataset, "gqa": gqa_clip_data.GQAFineTuneDataset,
               "nlvr": nlvr_clip_data.NLVRFineTuneDataset, "caption": caption_clip_data.COCOCaptionFineTuneDataset,
               "tvqa": tvqa_data.TVQAFineTuneDataset, "how2qa": how2qa_data.How2QAFineTuneDataset,
               "tvc": tvc_data.TVCFineTuneDataset, "yc2c": yc2c_data.YC2CFineTuneDataset}
    return classes, tvqa_data.resize, multitask_data.MultitaskLoader


def _q(t):
    """Values on a 1/64 grid: exact in fp32 and in the JSON text."""
    return torch.round(t * 64) / 64


def make_entries(task: str, B: int, seed: int, with_target: bool = True, V_L: int = 5, feat: int = 6):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(B):
        n_in = int(torch.randint(1, 8, (1,), generator=g))
        n_tg = int(torch.randint(1, 5, (1,), generator=g))
        e = {"input_ids": torch.randint(3, 99, (n_in,), generator=g).tolist(), "input_length": n_in,
             "sent": f"{task} sentence {i}", "question_id": 100 * seed + i}
        shape = (2, V_L) if task == "nlvr" else (V_L,)
        e["vis_feats"] = _q(torch.randn(*shape, feat, generator=g)).tolist()
        e["boxes"] = _q(torch.rand(*shape, 4, generator=g)).tolist()
        if with_target:
            e["target_ids"] = torch.randint(3, 99, (n_tg,), generator=g).tolist()
            e["target_length"] = n_tg
            e["answer"] = f"answer {i}"
        if task in ("vqa", "gqa"):
            e["all_answers"] = [f"answer {i}", "other"]
            e["label"] = {f"answer {i}": 1.0}
            if with_target:
                e["score"] = float(torch.randint(0, 4, (1,), generator=g)) / 4
        if task == "nlvr":
            e["label"] = int(i % 2) if with_target else None
        if task == "caption":
            e["n_boxes"] = V_L
            e["img_id"] = f"img{i}"
            e["input_text"] = "describe image with tags:"
            e["targets"] = [f"caption {i} a", f"caption {i} b"]
            e.pop("sent"), e.pop("question_id"), e.pop("answer", None)
        if task in ("tvc", "yc2c"):
            e["video_id"] = f"vid{i}"
        if task == "tvc":
            e["ts"] = f"{i}.0-{i + 3}.5"
        out.append(e)
    return out


def to_reference_entries(entries, args):
    """JSON entry -> what the reference's ``__getitem__`` returns (tensors, the shared ``args``)."""
    out = []
    for e in entries:
        r = dict(e)
        r["args"] = args
        r["input_ids"] = torch.LongTensor(e["input_ids"])
        if "target_ids" in e:
            r["target_ids"] = torch.LongTensor(e["target_ids"])
        r["vis_feats"] = torch.tensor(e["vis_feats"], dtype=torch.float32)
        r["boxes"] = torch.tensor(e["boxes"], dtype=torch.float32)
        out.append(r)
    return out


def jsonable(batch):
    out = {}
    for k, v in batch.items():
        if k == "args":
            continue
        if torch.is_tensor(v):
            out[k] = {"dtype": str(v.dtype).replace("torch.", ""), "shape": list(v.shape), "data": v.reshape(-1).tolist()}
        else:
            out[k] = v
    return out


class _FakeSampler:
    def set_epoch(self, epoch):
        self.epoch = epoch


class _FakeLoader:
    def __init__(self, task, n):
        self.task, self.n, self.sampler = task, n, _FakeSampler()

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter([{"task": self.task, "index": i} for i in range(self.n)])


def main():
    classes, resize, MultitaskLoader = _import_data_modules()
    args = types.SimpleNamespace(use_vision=True, no_prefix=False)
    cases = []
    for ti, task in enumerate(IMAGE_TASKS + VIDEO_TASKS):
        for with_target in (True, False):
            entries = make_entries(task, B=3 if with_target else 2, seed=10 * ti + int(with_target), with_target=with_target)
            fake_self = types.SimpleNamespace(tokenizer=types.SimpleNamespace(pad_token_id=PAD), args=args)
            batch = classes[task].collate_fn(fake_self, to_reference_entries(entries, args))
            cases.append({"task": task, "pad_token_id": PAD, "entries": entries, "batch": jsonable(batch)})
    resize_cases = []
    g = torch.Generator().manual_seed(7)
    for L, length in ((5, 8), (8, 8), (16, 8), (13, 8), (9, 4), (1, 3)):
        x = _q(torch.randn(L, 4, generator=g))
        resize_cases.append({"length": length, "input": x.tolist(), "output": resize(x, length).tolist()})
    schedule_cases = []
    lens = {"vqa": 3, "gqa": 5, "nlvr": 1, "caption": 2}
    for sampling, n_batches in (("roundrobin", None), ("balanced", None), ("balanced", 2)):
        for shuffle in (True, False):
            ml = MultitaskLoader([_FakeLoader(t, n) for t, n in lens.items()], shuffle=shuffle, sampling=sampling,
                                 n_batches=n_batches, verbose=False)
            for epoch in (0, 1, 5):
                ml.set_epoch(epoch)
                n = len(ml)
                if sampling == "roundrobin":
                    order = [(b["task"], b["index"]) for b in ml]
                else:   # 'balanced' asks a loader for more batches than it has: only the task order is defined
                    order = [(t, None) for t in reversed(ml.epoch_tasks)]
                schedule_cases.append({"lens": lens, "sampling": sampling, "n_batches": n_batches, "shuffle": shuffle,
                                       "epoch": epoch, "len": n, "order": order})
    path = os.path.join(HERE, "collate_cases.json")
    with open(path, "w") as f:
        json.dump({"collate": cases, "resize": resize_cases, "schedule": schedule_cases}, f)
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "collate,", len(resize_cases), "resize,",
          len(schedule_cases), "schedule cases")


if __name__ == "__main__":
    random.seed(0)
    np.random.seed(0)
    main()
