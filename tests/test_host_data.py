"""SURVEY section 8 row f-4, data / wire formats: ``vlpet_b200.host.data`` against ``tests/golden/collate_cases.json`` -- the
reference's own ``collate_fn`` of all eight tasks, the video loaders' ``resize`` and ``MultitaskLoader``'s task order, run
unmodified on seeded entries by ``tests/golden/make_golden_collate.py``.  Integer and index work: bit-exact.  CPU only."""
import json
import os

import pytest
import torch

from tests.helpers import GOLDEN


@pytest.fixture(scope="module")
def D():
    try:
        import vlpet_b200.host.data as D_
    except Exception as e:                       # the package needs libvlpet.so (built in-tree by build())
        pytest.skip(f"vlpet_b200 not importable: {e}")
    return D_


@pytest.fixture(scope="module")
def cases():
    with open(os.path.join(GOLDEN, "collate_cases.json")) as f:
        return json.load(f)


def _tensor(spec):
    return torch.tensor(spec["data"], dtype=getattr(torch, spec["dtype"])).reshape(spec["shape"])


def test_collate_matches_reference_collate_fn(D, cases):
    seen = set()
    for c in cases["collate"]:
        got = D.collate(c["task"], c["entries"], c["pad_token_id"])
        want = c["batch"]
        assert set(got) == set(want), (c["task"], sorted(set(got) ^ set(want)))
        for k, v in want.items():
            if isinstance(v, dict) and "dtype" in v:
                w = _tensor(v)
                assert got[k].dtype == w.dtype and got[k].shape == w.shape, (c["task"], k, got[k].dtype, got[k].shape)
                assert torch.equal(got[k], w), (c["task"], k)
            else:
                assert got[k] == v, (c["task"], k)
        seen.add((c["task"], "target_ids" in want))
    assert seen == {(t, f) for t in D.IMAGE_TASKS + D.VIDEO_TASKS for f in (True, False)}


def test_collate_accepts_tensors_and_arrays_and_refuses_ragged_grids(D, cases):
    import numpy as np
    c = cases["collate"][0]
    as_t = [dict(e, input_ids=torch.tensor(e["input_ids"]), vis_feats=np.asarray(e["vis_feats"], dtype=np.float32),
                 boxes=torch.tensor(e["boxes"])) for e in c["entries"]]
    a, b = D.collate("vqa", as_t, 1), D.collate("vqa", c["entries"], 1)
    assert all(torch.equal(a[k], b[k]) for k in ("input_ids", "target_ids", "vis_feats", "boxes", "scores"))
    bad = [dict(e) for e in c["entries"]]
    bad[1]["vis_feats"] = bad[1]["vis_feats"][:-1]
    with pytest.raises(ValueError):
        D.collate("vqa", bad, 1)
    with pytest.raises(ValueError):
        D.collate("vqa", [], 1)
    with pytest.raises(ValueError):
        D.collate("okvqa", c["entries"], 1)


def test_collate_feeds_the_host_model_schema(D, cases):
    """The batch keys ``train_step`` reads (host/synthetic.py's schema) are a subset of what collate produces."""
    for c in cases["collate"]:
        if "target_ids" not in c["batch"]:
            continue
        got = D.collate(c["task"], c["entries"], c["pad_token_id"])
        need = {"task", "input_ids", "vis_feats", "boxes", "target_ids"} | ({"scores"} if c["task"] in ("vqa", "gqa") else set())
        assert need <= set(got)
        assert int((got["target_ids"] == c["pad_token_id"]).sum()) == 0 and got["target_ids"].min() >= -100


def test_resize_frames_matches_reference_resize(D, cases):
    for c in cases["resize"]:
        x = torch.tensor(c["input"], dtype=torch.float32)
        y = D.resize_frames(x, c["length"])
        assert y.shape == (c["length"], x.shape[1])
        assert torch.equal(y, torch.tensor(c["output"], dtype=torch.float32)), (x.shape, c["length"])


class _Loader:
    def __init__(self, task, n):
        self.task, self.n = task, n

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter([(self.task, i) for i in range(self.n)])


def test_multitask_loader_order_matches_reference(D, cases):
    for c in cases["schedule"]:
        ml = D.MultitaskLoader([_Loader(t, n) for t, n in c["lens"].items()], shuffle=c["shuffle"], sampling=c["sampling"],
                               n_batches=c["n_batches"])
        ml.set_epoch(c["epoch"])
        assert len(ml) == c["len"]
        if c["sampling"] == "roundrobin":
            assert [list(b) for b in ml] == c["order"], c
            assert len(ml) == 0
        else:
            assert [[t, None] for t in reversed(ml.epoch_tasks)] == c["order"], c


def test_task_batches_split_like_distributed_sampler(D, cases):
    from torch.utils.data.distributed import DistributedSampler
    entries = [dict(cases["collate"][0]["entries"][i % 3], question_id=i) for i in range(11)]
    for world in (1, 2, 3):
        for shuffle in (False, True):
            for epoch in (0, 2):
                got = []
                for rank in range(world):
                    tb = D.TaskBatches("vqa", entries, batch_size=2, pad_token_id=1, rank=rank, world=world, shuffle=shuffle, seed=5)
                    tb.sampler.set_epoch(epoch)
                    ref = DistributedSampler(entries, num_replicas=world, rank=rank, shuffle=shuffle, seed=5)
                    ref.set_epoch(epoch)
                    ids = [q for b in tb for q in b["question_ids"]]
                    assert ids == list(ref), (world, rank, shuffle, epoch)
                    assert len(tb) == -(-len(ids) // 2)
                    got += ids
                assert set(got) == set(range(11))
    ml = D.MultitaskLoader([D.TaskBatches("vqa", entries, 4, 1), D.TaskBatches("gqa", entries[:5], 4, 1)])
    tasks = [b["task"] for b in ml]
    assert sorted(tasks) == ["gqa", "gqa", "vqa", "vqa", "vqa"]


def test_collated_batch_reproduces_reference_loss_cpu(D):
    """Entries -> collate -> host.VLBart.train_step gives the loss the reference's VLBart gave on the same samples
    (tests/golden/vlbart_tiny_large.npz): the wire format and the caller of the hot path agree end to end."""
    import numpy as np
    import vlpet_b200.host as H
    from oracle.eager_ref import use_eager_pet
    from tests.test_host_model import _batch, _cfg, _load, _load_state
    z = _load("large")
    model = use_eager_pet(H.VLBart(_cfg(H, "large")).double().eval())
    _load_state(model, z, torch.float64)
    pad = model.config.pad_token_id
    for task in ("vqa", "nlvr"):
        want = _batch(z, task, torch.float64)
        entries = []
        for i in range(want["input_ids"].shape[0]):
            ids, tgt = want["input_ids"][i], want["target_ids"][i]
            e = {"input_ids": ids[ids != pad].tolist(), "target_ids": tgt[tgt != -100].tolist(), "sent": "", "question_id": i,
                 "vis_feats": z[f"{task}/vis_feats"][i], "boxes": z[f"{task}/boxes"][i], "label": None}
            if "scores" in want:
                e["score"] = float(want["scores"][i])
            entries.append(e)
        got = D.collate(task, entries, pad)
        for k in ("input_ids", "target_ids"):
            assert torch.equal(got[k], want[k]), (task, k)
        assert np.array_equal(got["vis_feats"].numpy(), z[f"{task}/vis_feats"].astype(np.float32))
        got["vis_feats"], got["boxes"] = want["vis_feats"], want["boxes"]          # fp64 model: keep the fixture's precision
        if "scores" in want:
            assert torch.equal(got["scores"], want["scores"].float())
            got["scores"] = want["scores"]
        loss = model.train_step(got)["loss"]
        assert abs(loss.item() - float(z[f"{task}/loss"])) < 1e-9, task


def test_fit_and_predict_over_a_multitask_loader_cpu(D):
    """entries -> TaskBatches -> MultitaskLoader -> PetTrainer.fit / predict (the epoch loop of multitask.py:189-345 and the
    predict half of its ``*_evaluate``): task order and counts follow the loader, the per-task mean losses equal a hand-written
    loop over the same loader, predictions equal ``test_step`` called batch by batch.  fp64 on CPU with the oracle on the PET
    sites; the fused AdamW kernel is CUDA-only, so the test steps with plain SGD on the flat bucket."""
    import vlpet_b200.host as H
    from oracle.eager_ref import use_eager_pet
    from tests.test_host_model import _cfg

    class SgdTrainer(H.PetTrainer):
        def optimizer_step(self):
            lr = H.linear_warmup_lr(self.step_idx, self.total_steps, self.warmup_ratio, self.lr)
            self.bucket.flat_param.add_(self.bucket.flat_grad, alpha=-lr)
            self.step_idx += 1
            self.opt_steps += 1

    cfg = _cfg(H, "large")
    g = torch.Generator().manual_seed(11)

    def entry(task, i, target=True):
        shape = (2, cfg.n_boxes) if task == "nlvr" else (cfg.n_boxes,)
        e = {"input_ids": torch.randint(3, 290, (int(torch.randint(2, 7, (1,), generator=g)),), generator=g).tolist(),
             "vis_feats": torch.randn(*shape, cfg.feat_dim, generator=g), "boxes": torch.zeros(*shape, 4),
             "sent": f"s{i}", "question_id": f"{task}{i}", "label": None}
        if target:
            e.update(target_ids=torch.randint(3, 290, (int(torch.randint(1, 4, (1,), generator=g)),), generator=g).tolist(),
                     answer=f"a{i}", score=1.0)
        return e

    data = {"vqa": [entry("vqa", i) for i in range(7)], "nlvr": [entry("nlvr", i) for i in range(4)]}

    def loader():
        return D.MultitaskLoader([D.TaskBatches(t, es, 3, cfg.pad_token_id) for t, es in data.items()])

    def trainer():
        torch.manual_seed(0)
        model = use_eager_pet(H.VLBart(cfg).double())
        return SgdTrainer(model, cfg, "cpu", lr=1e-2, total_steps=10, compute_dtype=torch.float64)

    t1 = trainer()
    hist = t1.fit(loader(), epochs=2)
    assert [h["epoch"] for h in hist] == [0, 1] and all(h["task_counter"] == {"vqa": 3, "nlvr": 2} and h["steps"] == 5 for h in hist)
    assert t1.step_idx == 10 and hist[1]["loss"] < hist[0]["loss"]
    t2 = trainer()
    ld = loader()
    for epoch in range(2):
        ld.set_epoch(epoch)
        sums = {"vqa": 0.0, "nlvr": 0.0}
        for b in ld:
            sums[b["task"]] += float(t2.train_step(b))
        assert abs(sums["vqa"] / 3 - hist[epoch]["task_loss"]["vqa"]) < 1e-12
        assert abs(sums["nlvr"] / 2 - hist[epoch]["task_loss"]["nlvr"]) < 1e-12
    assert torch.equal(t1.bucket.flat_param, t2.bucket.flat_param)
    seen = []
    t1.fit(loader(), epochs=1, start_epoch=2, on_epoch_end=lambda tr, rec: seen.append((tr is t1, rec["epoch"])))
    assert seen == [(True, 2)]

    test = D.TaskBatches("vqa", [entry("vqa", 100 + i, target=False) for i in range(4)], 3, cfg.pad_token_id)
    pred = t1.predict(test, max_length=5)
    assert set(pred) == {"vqa"} and sorted(pred["vqa"]) == [f"vqa{100 + i}" for i in range(4)]
    t1.model.eval()
    for b in test:
        tok = t1.model.test_step(b, max_length=5)["token_ids"].tolist()
        assert [pred["vqa"][q] for q in b["question_ids"]] == tok


def test_synthetic_cycle_ratios_and_shards():
    """host/synthetic.py: the per-task batch ratios of multitask.py:682-695 (vqa b, gqa b*100/60, nlvr b*20/60, caption b*50/60;
    the video tasks unscaled, multitask_video.py:740-743) and the per-rank shards of one global cycle."""
    import vlpet_b200.host as H
    assert H.task_batch_sizes(300) == {"vqa": 300, "gqa": 500, "nlvr": 100, "caption": 250}
    assert H.task_batch_sizes(50, H.VIDEO_TASKS) == {t: 50 for t in H.VIDEO_TASKS}
    tasks = ["vqa", "gqa", "nlvr", "caption"]
    full = H.multitask_cycle(6, tasks, feat_dim=8, seed=3)
    assert [b["task"] for b in full] == tasks and [b["input_ids"].shape[0] for b in full] == [6, 10, 2, 5]
    assert full[2]["vis_feats"].shape == (2, 2, 49, 8) and full[0]["vis_feats"].shape == (6, 49, 8)
    assert all(int((b["boxes"] != 0).sum()) == 0 for b in full)          # CLIP grid features carry no boxes (vqa_clip_data.py:198)
    for world in (2, 4):
        parts = [H.multitask_cycle(6, tasks, feat_dim=8, seed=3, rank=r, world=world) for r in range(world)]
        for i, b in enumerate(full):
            for k, v in b.items():
                if torch.is_tensor(v):
                    assert torch.equal(torch.cat([p[i][k] for p in parts]), v), (world, b["task"], k)
    assert H.batch_nbytes(full[0]) == sum(v.numel() * v.element_size() for v in full[0].values() if torch.is_tensor(v))
    again = H.multitask_cycle(6, tasks, feat_dim=8, seed=3)
    assert all(torch.equal(a["input_ids"], b["input_ids"]) for a, b in zip(full, again))
