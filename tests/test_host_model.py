"""The caller of the hot path: ``vlpet_b200.host.VLBart`` against golden vectors produced by the
reference's own ``VLBart`` (tests/golden/make_golden_vlbart.py) -- state_dict loaded key for key, loss and every
trainable gradient compared.

CPU (not gpu): the PET sites run the eager restatement of oracle/eager_ref.py (fp64, tol 1e-9) -- this checks the
host plumbing (embeddings, attention, LayerNorms, decoder, loss shaping, parameter names).
GPU: the same model with the CUDA kernels on the PET sites, fp32 (tol 2e-4 on gradients: fp32 accumulation order
over a 4-layer model) -- this is the drop-in claim end to end.
"""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, rel

CASES = ["large", "small"]


def _load(gate):
    z = np.load(os.path.join(GOLDEN, f"vlbart_tiny_{gate}.npz"), allow_pickle=False)
    return z


def _cfg(H, gate, **kw):
    flags = dict(use_encoder_adapter_gating_large_x_lowrank=(gate == "large"),
                 use_encoder_adapter_gating_small_xy_cat=(gate == "small"), dropout=0.0, attention_dropout=0.0,
                 activation_dropout=0.0)
    flags.update(kw)
    return H.tiny_test_config(**flags)


def _batch(z, task, dtype):
    b = {"task": task, "input_ids": torch.tensor(z[f"{task}/input_ids"]), "target_ids": torch.tensor(z[f"{task}/target_ids"]),
         "vis_feats": torch.tensor(z[f"{task}/vis_feats"]).to(dtype), "boxes": torch.tensor(z[f"{task}/boxes"]).to(dtype)}
    if f"{task}/scores" in z.files:
        b["scores"] = torch.tensor(z[f"{task}/scores"]).to(dtype)
    return b


def _load_state(model, z, dtype):
    keys = [str(k) for k in z["meta_state_keys"]]
    sd = {k: torch.tensor(z["sd/" + k]).to(dtype) if z["sd/" + k].dtype.kind == "f" else torch.tensor(z["sd/" + k])
          for k in keys}
    ours = model.state_dict()
    assert set(ours.keys()) == set(keys), (sorted(set(ours) - set(keys))[:5], sorted(set(keys) - set(ours))[:5])
    model.load_state_dict(sd, strict=True)
    return keys


@pytest.fixture(scope="module")
def H():
    try:
        import vlpet_b200.host as H_
    except Exception as e:                       # the package needs libvlpet.so (built in-tree by build())
        pytest.skip(f"vlpet_b200 not importable: {e}")
    return H_


@pytest.mark.parametrize("gate", CASES)
def test_state_dict_names_and_trainable_set_match_reference(H, gate):
    z = _load(gate)
    model = H.VLBart(_cfg(H, gate))
    _load_state(model, z, torch.float32)
    assert sorted(H.trainable_names(model, model.config)) == sorted(str(n) for n in z["meta_trainable"])


@pytest.mark.parametrize("gate", CASES)
def test_host_model_matches_reference_vlbart_cpu(H, gate):
    from oracle.eager_ref import use_eager_pet
    z = _load(gate)
    model = use_eager_pet(H.VLBart(_cfg(H, gate)).double().eval())
    _load_state(model, z, torch.float64)
    names = [str(n) for n in z["meta_trainable"]]
    params = dict(model.named_parameters())
    for task in ("vqa", "nlvr"):
        model.zero_grad()
        loss = model.train_step(_batch(z, task, torch.float64))["loss"]
        loss.backward()
        assert abs(loss.item() - float(z[f"{task}/loss"])) < 1e-10
        for n in names:
            assert rel(params[n].grad.numpy(), z[f"{task}/grad/{n}"]) < 1e-9, (task, n)


def _freeze_backbone(H, model):
    """What PetTrainer does before a step (trainer_base.py:268-270, 308-542): only the PET parameters stay trainable.  The
    host models fuse frozen projections (q/k/v of a self-attention, the cross-attention keys of all decoder layers) only then."""
    names = set(H.trainable_names(model, model.config))
    for n, p in model.named_parameters():
        p.requires_grad_(n in names)


@pytest.mark.parametrize("arch", ["bart", "t5"])
def test_frozen_backbone_fusions_match_reference_cpu(H, arch):
    """With the backbone frozen the host models run their fused projections (one GEMM for q/k/v of a frozen self-attention,
    one GEMM for the cross-attention keys of all decoder layers): loss and every trainable gradient still equal the
    reference's (the goldens of make_golden_vlbart.py / make_golden_vlt5.py)."""
    from oracle.eager_ref import use_eager_pet
    if arch == "bart":
        z = _load("large")
        model = use_eager_pet(H.VLBart(_cfg(H, "large")).double().eval())
        tl, tg = 1e-10, 1e-9
    else:
        z = _load_t5()
        model = use_eager_pet(H.VLT5(_t5_cfg(H)).double().eval())
        tl, tg = 1e-7, 2e-6
    _load_state(model, z, torch.float64)
    _freeze_backbone(H, model)
    dec = model.model.decoder if arch == "bart" else model.decoder
    names = [str(n) for n in z["meta_trainable"]]
    params = dict(model.named_parameters())
    for task in ("vqa", "nlvr"):
        model.zero_grad()
        loss = model.train_step(_batch(z, task, torch.float64))["loss"]
        loss.backward()
        assert dec._kcat is not None, "the batched cross-attention key projection did not run"
        assert abs(loss.item() - float(z[f"{task}/loss"])) < tl
        for n in names:
            assert rel(params[n].grad.numpy(), z[f"{task}/grad/{n}"]) < tg, (task, n)


def _generate_case(H, dtype, device, eager, arch="bart"):
    z = np.load(os.path.join(GOLDEN, f"vl{arch}_tiny_generate.npz"), allow_pickle=False)
    model = H.VLBart(_cfg(H, "large")) if arch == "bart" else H.VLT5(H.tiny_t5_test_config(dropout_rate=0.0, dropout=0.0))
    model = model.double() if dtype == torch.float64 else model
    if eager:
        from oracle.eager_ref import use_eager_pet
        model = use_eager_pet(model)
    model.eval()
    _load_state(model, z, dtype)
    model.to(device)
    ids = torch.tensor(z["vqa/input_ids"]).to(device)
    vis = (torch.tensor(z["vqa/vis_feats"]).to(device=device, dtype=dtype), torch.tensor(z["vqa/boxes"]).to(device=device, dtype=dtype))
    bias = torch.tensor(z["vqa/logit_bias"]).to(device)
    proc = lambda step, tokens, scores: scores + bias[:, step].to(scores.dtype)  # noqa: E731
    return z, model, ids, vis, proc


def test_generate_matches_reference_cached_decode_cpu(H):
    """greedy decoding through the KV cache (cross-attention keys / values -- the values through the value
    parallel adapter -- formed once) against the reference's own cached decode path
    (tests/golden/make_golden_generate.py: VLBart.forward with past_key_values, src/modeling_bart.py:1522-1602), host logic
    in fp64 with the eager PET restatement: same tokens, every step's logits to 1e-9; and the cached steps reproduce the
    model's own full teacher-forced pass."""
    z, model, ids, vis, proc = _generate_case(H, torch.float64, "cpu", eager=True)
    tokens, logits = model.generate(ids, vis, task="vqa", max_length=int(z["meta_max_length"]), min_length=int(z["meta_min_length"]),
                                    logits_processor=proc, return_step_logits=True)
    assert tokens.tolist() == z["vqa/tokens"].tolist()
    assert rel(logits.numpy(), z["vqa/step_logits"]) < 1e-9
    with torch.no_grad():
        h = model.model(ids, vis, tokens[:, :-1], task="vqa")
        full = model._logits(h)
    assert rel(full.numpy(), logits.numpy()) < 1e-9
    assert model.test_step({"task": "vqa", "input_ids": ids, "vis_feats": vis[0], "boxes": vis[1]}, max_length=4, min_length=4)["token_ids"].shape == (3, 4)


@pytest.mark.parametrize("arch", ["bart", "t5"])
def test_beam_search_over_the_cached_step_cpu(H, arch):
    """host/generation.py (the caption task's generate(num_beams=5)): (1) one beam is greedy decoding; (2) on a vocabulary
    restricted to 3 tokens and 4 generated positions, num_beams = 27 is exhaustive, so its result must be the sequence an
    enumeration of all 81 continuations scores highest under a full teacher-forced pass; (3) the sum of log-probabilities the
    search reports for a 4-beam run equals a teacher-forced re-scoring of its output -- the self-attention caches followed the
    beam re-ordering."""
    z, model, ids, vis, proc = _generate_case(H, torch.float64, "cpu", eager=True, arch=arch)
    L, mn = int(z["meta_max_length"]), int(z["meta_min_length"])
    eos = model.config.eos_token_id

    def full_logp(tokens):                                   # log-probabilities of tokens[:, 1:] under one teacher-forced pass
        B = tokens.shape[0]
        rep = torch.arange(ids.shape[0]).repeat_interleave(B // ids.shape[0])
        with torch.no_grad():
            if arch == "bart":
                h = model.model(ids[rep], tuple(v[rep] for v in vis), tokens[:, :-1], task="vqa")
                lg = model._logits(h)
            else:
                enc, mask = model.encoder(ids[rep], tuple(v[rep] for v in vis), task="vqa")
                lg = model.lm_head(model.decoder(tokens[:, :-1], enc, encoder_mask=mask, task="vqa") * (model.model_dim ** -0.5))
        return torch.log_softmax(lg.double(), -1)

    # (1) one beam == greedy (the processor adds the same bias to logits and to log-probabilities: argmax unchanged)
    greedy = model.generate(ids, vis, task="vqa", max_length=L, min_length=L, logits_processor=proc)
    one = model.generate(ids, vis, task="vqa", max_length=L, min_length=L, num_beams=1, logits_processor=proc)
    beam1 = model.generate(ids[:1], (vis[0][:1], vis[1][:1]), task="vqa", max_length=L, min_length=L, num_beams=2,
                           logits_processor=lambda s_, t_, sc: sc, length_penalty=1.0)
    assert greedy.tolist() == one.tolist() and beam1.shape == (1, L)
    # (2) exhaustive on 3 allowed tokens, 4 generated positions
    allowed = torch.tensor([5, 17, 123])
    Lx = 5

    def restrict(step, tokens, scores):
        m = torch.full_like(scores, -float("inf"))
        m[:, allowed] = 0.0
        return scores + m

    out, rep_scores = model.generate(ids, vis, task="vqa", max_length=Lx, min_length=Lx, num_beams=27, logits_processor=restrict,
                                     return_scores=True)
    import itertools
    cont = torch.tensor(list(itertools.product(allowed.tolist(), repeat=Lx - 1)))            # [81, 4]
    B = ids.shape[0]
    start = torch.full((cont.shape[0], 1), model.config.decoder_start_token_id, dtype=torch.long)
    seqs = torch.cat([start, cont], 1).repeat(B, 1)                                          # sample-major
    lp = full_logp(seqs)            # (the processor masks log-probabilities, as in HF: allowed tokens keep their full-vocabulary values)
    tot = lp.gather(-1, seqs[:, 1:, None]).squeeze(-1).sum(-1).view(B, -1)
    best = tot.argmax(-1)
    for b in range(B):
        assert out[b].tolist() == seqs[b * cont.shape[0] + int(best[b])].tolist(), b
        assert abs(float(rep_scores[b]) - float(tot[b, best[b]])) < 1e-6 * max(1.0, abs(float(tot[b, best[b]])))
    # (3) reported score == teacher-forced re-scoring (4 beams, full vocabulary, EOS suppressed so all hypotheses have length L)
    out4, sc4 = model.generate(ids, vis, task="vqa", max_length=L, min_length=L, num_beams=4, return_scores=True)
    lp4 = full_logp(out4)
    lp4[..., eos] = -float("inf")                                                             # (what min_length did at every step)
    re = lp4.gather(-1, out4[:, 1:, None]).squeeze(-1).sum(-1)
    assert torch.allclose(re, sc4, rtol=1e-6, atol=1e-6), (re, sc4)
    g = model.generate(ids, vis, task="vqa", max_length=L, min_length=L)                      # plain greedy, no processor
    lpg = full_logp(g)
    assert bool((sc4 >= lpg.gather(-1, g[:, 1:, None]).squeeze(-1).sum(-1) - 1e-9).all())     # 4 beams found at least the greedy path here


def test_trainer_checkpoint_resume_round_trip_cpu(H, tmp_path):
    """PetTrainer.save / load (trainer_base.py:764-781): a PET-only ``<name>.pth`` under the reference's key names + the
    optimizer state a resume needs.  A second trainer on a freshly initialised model loads them: same parameters (masters still
    pinned in the flat bucket), same AdamW moments, same update / LR counters, same loss; the key report is clean."""
    from oracle.eager_ref import use_eager_pet
    z = _load("large")

    def make(seed):
        torch.manual_seed(seed)
        model = use_eager_pet(H.VLBart(_cfg(H, "large")).double().eval())
        return model, H.PetTrainer(model, model.config, "cpu", compute_dtype=torch.float64)

    m1, t1 = make(0)
    _load_state(m1, z, torch.float64)
    batch = _batch(z, "vqa", torch.float64)
    loss1 = float(t1.forward_backward(batch))
    g = torch.Generator().manual_seed(3)
    t1.bucket.exp_avg.copy_(torch.randn(t1.bucket.numel, generator=g, dtype=torch.float64))
    t1.bucket.exp_avg_sq.copy_(torch.rand(t1.bucket.numel, generator=g, dtype=torch.float64))
    t1.opt_steps, _ = 17, t1.set_step(23)
    path = str(tmp_path / "LAST")
    t1.save(path)
    saved = torch.load(path + ".pth")
    assert set(saved) == set(t1.bucket.names) and all(v.dtype == torch.float64 for v in saved.values())
    m2, t2 = make(1)                                             # different init: everything trainable must come from the file
    for n, p in m2.named_parameters():                           # the frozen backbone comes from the pretrained weights, as in the reference
        if n not in saved:
            p.data.copy_(dict(m1.named_parameters())[n].data)
    res = t2.load(path)
    assert not res.unexpected_keys and set(res.missing_keys).isdisjoint(saved)
    assert torch.equal(t2.bucket.flat_param, t1.bucket.flat_param)
    for p, o in zip(t2.bucket.params, t2.bucket.offsets):        # still views of the bucket
        assert p.data_ptr() == t2.bucket.flat_param.data_ptr() + 8 * o
    assert torch.equal(t2.bucket.exp_avg, t1.bucket.exp_avg) and torch.equal(t2.bucket.exp_avg_sq, t1.bucket.exp_avg_sq)
    assert (t2.opt_steps, t2.step_idx) == (17, 23)
    assert abs(float(t2.forward_backward(batch)) - loss1) < 1e-12
    t1.save(path + "_full", full=True, with_optimizer=False)
    assert set(torch.load(path + "_full.pth")) == set(m1.state_dict())


def test_no_repeat_ngram_processor():
    """host.generation.NoRepeatNGram == HF's NoRepeatNGramLogitsProcessor semantics: a token completing an n-gram that already
    occurred is banned; nothing is banned before n - 1 tokens exist."""
    from vlpet_b200.host.generation import NoRepeatNGram, chain
    toks = torch.tensor([[7, 1, 2, 3, 1, 2], [7, 5, 5, 5, 5, 5], [7, 1, 2, 9, 4, 8]])
    sc = NoRepeatNGram(3)(5, toks, torch.zeros(3, 12))
    assert torch.isinf(sc[0, 3]) and torch.isinf(sc[0]).sum() == 1            # (1, 2) was followed by 3
    assert torch.isinf(sc[1, 5]) and torch.isinf(sc[1]).sum() == 1            # (5, 5) was followed by 5
    assert not torch.isinf(sc[2]).any()
    assert not torch.isinf(NoRepeatNGram(3)(0, toks[:, :1], torch.zeros(3, 12))).any()
    sc1 = chain(NoRepeatNGram(1), NoRepeatNGram(2))(5, toks, torch.zeros(3, 12))
    assert set(torch.isinf(sc1[0]).nonzero().flatten().tolist()) == {7, 1, 2, 3}   # unigrams: every token seen so far
    from transformers import NoRepeatNGramLogitsProcessor                      # the library's processor on random sequences
    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 3, 4):
        for cur_len in (1, 2, 3, 6, 11):
            rnd = torch.randint(0, 5, (6, cur_len), generator=g)
            want = NoRepeatNGramLogitsProcessor(n)(rnd, torch.zeros(6, 5))
            assert torch.equal(NoRepeatNGram(n)(cur_len - 1, rnd, torch.zeros(6, 5)), want), (n, cur_len)


@pytest.mark.gpu
def test_generate_with_cuda_vpa_matches_reference_cached_decode(H):
    """The same on the GPU in fp32: the value parallel adapter inside ``cross_kv`` is the K2 CUDA kernel (forward only, once per
    generation); tokens identical, step logits to 2e-4."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200 as V
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z, model, ids, vis, proc = _generate_case(H, torch.float32, "cuda", eager=False)
    n0 = V.launch_count()
    tokens, logits = model.generate(ids, vis, task="vqa", max_length=int(z["meta_max_length"]), min_length=int(z["meta_min_length"]),
                                    logits_processor=proc, return_step_logits=True)
    assert V.launch_count() - n0 >= 2 + 2 * 2 + 1, "PET sites did not run the CUDA kernels"   # K2 per decoder layer, K1 x 2 per encoder layer, K3
    assert tokens.cpu().tolist() == z["vqa/tokens"].tolist()
    assert rel(logits.double().cpu().numpy(), z["vqa/step_logits"]) < 2e-4


def test_bart_base_trainable_count_is_the_reference_checksum(H):
    """BART-base + VL-PET-large r=96: 6 052 416 trainable parameters = the 4.16 % of the reference README
    (README.md:360; reproduced by instantiating the reference model)."""
    cfg = H.bart_base_vlpet_large()
    with torch.device("meta"):
        model = H.VLBart(cfg)
    names = set(H.trainable_names(model, cfg))
    n = sum(p.numel() for k, p in model.named_parameters() if k in names)
    assert n == 6052416


def test_trainable_share_matches_the_published_table(H):
    """The 'Trainable Params (%)' column of the reference's results table (README.md:357-360: small / middleX / middleY 2.98,
    large 4.16; all at r = 96) from the host model's parameter names and the trainer's unfreeze rules."""
    def share(cfg):
        with torch.device("meta"):
            model = H.VLBart(cfg)
        names = set(H.trainable_names(model, cfg))
        return 100.0 * sum(p.numel() for k, p in model.named_parameters() if k in names) / sum(p.numel() for p in model.parameters())

    small = H.bart_base_vlpet_small(r=96, dec_r=96)
    middle = [small.clone(use_encoder_adapter_gating_small_xy_cat=False, **{f: True})
              for f in ("use_encoder_adapter_gating_middle_xy_add", "use_encoder_adapter_gating_middle_ia3_add")]
    for cfg, want in ((H.bart_base_vlpet_large(), 4.16), (small, 2.98), (middle[0], 2.98), (middle[1], 2.98)):
        assert round(share(cfg), 2) == want


def test_task_batch_ratios_follow_the_reference():
    from vlpet_b200.host import task_batch_sizes
    assert task_batch_sizes(500) == {"vqa": 500, "gqa": 833, "nlvr": 166, "caption": 416}
    assert task_batch_sizes(300) == {"vqa": 300, "gqa": 500, "nlvr": 100, "caption": 250}


def test_linear_warmup_schedule():
    from vlpet_b200.host import linear_warmup_lr
    assert linear_warmup_lr(0, 100, 0.1, 1e-3) == 0.0
    assert abs(linear_warmup_lr(5, 100, 0.1, 1e-3) - 5e-4) < 1e-12
    assert abs(linear_warmup_lr(10, 100, 0.1, 1e-3) - 1e-3) < 1e-12
    assert abs(linear_warmup_lr(55, 100, 0.1, 1e-3) - 5e-4) < 1e-12
    # the library scheduler the reference builds (trainer_base.py:722-730), stepped once per update
    from transformers import get_linear_schedule_with_warmup
    for total, ratio in ((37, 0.1), (100, 0.05), (10, 0.0), (64, 0.5)):
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=3e-4)
        sched = get_linear_schedule_with_warmup(opt, int(total * ratio), total)
        for step in range(total + 3):
            assert abs(linear_warmup_lr(step, total, ratio, 3e-4) - sched.get_last_lr()[0]) < 1e-15, (total, ratio, step)
            opt.step()
            sched.step()


@pytest.mark.gpu
@pytest.mark.parametrize("gate", CASES)
def test_host_model_with_cuda_pet_matches_reference_vlbart(H, gate):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200 as V
    z = _load(gate)
    # TF32 off: the frozen backbone must be fp32-exact for a 1e-4-level comparison
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = H.VLBart(_cfg(H, gate)).eval()
    _load_state(model, z, torch.float32)
    model.cuda()
    names = [str(n) for n in z["meta_trainable"]]
    params = dict(model.named_parameters())
    for n in names:
        params[n].requires_grad_(True)
    n0 = V.launch_count()
    for task in ("vqa", "nlvr"):
        model.zero_grad()
        loss = model.train_step(_batch(z, task, torch.float32))["loss"]
        loss.backward()
        assert abs(loss.item() - float(z[f"{task}/loss"])) < 2e-5 * abs(float(z[f"{task}/loss"]))
        for n in names:
            assert rel(params[n].grad.double().cpu().numpy(), z[f"{task}/grad/{n}"]) < 2e-4, (task, n)
    assert V.launch_count() - n0 >= 2 * (2 * 2 * 2 + 2 + 1), "PET sites did not run the CUDA kernels"


@pytest.mark.gpu
def test_trainer_bucket_gradients_match_reference_vlbart(H):
    """Same comparison through the PET-only trainer: gradients accumulated by the kernels STRAIGHT into the flat
    all-reduce bucket (functional.set_direct_grad_accumulation) must equal the reference VLBart gradients."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200.functional as F_
    z = _load("large")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = H.VLBart(_cfg(H, "large")).eval()
    _load_state(model, z, torch.float32)
    tr = H.PetTrainer(model, model.config, "cuda", compute_dtype=torch.float32)
    try:
        assert tr._direct and not F_._direct_grads      # direct accumulation is scoped to the trainer's own steps
        for task in ("vqa", "nlvr"):
            loss = tr.forward_backward(_batch(z, task, torch.float32))
            assert abs(loss.item() - float(z[f"{task}/loss"])) < 2e-5 * abs(float(z[f"{task}/loss"]))
            flat = tr.bucket.flat_grad.double().cpu().numpy()
            for n, p, o in zip(tr.bucket.names, tr.bucket.params, tr.bucket.offsets):
                got = flat[o:o + p.numel()]
                assert rel(got, z[f"{task}/grad/{n}"].reshape(-1)) < 2e-4, (task, n)
    finally:
        F_.set_direct_grad_accumulation(False)


# ------------------------------------------------------------------------------------------ T5 (BASELINE config 3)
def _load_t5():
    return np.load(os.path.join(GOLDEN, "vlt5_tiny_large.npz"), allow_pickle=False)


def _t5_cfg(H, **kw):
    return H.tiny_t5_test_config(dropout_rate=0.0, dropout=0.0, **kw)


def test_t5_state_dict_names_and_trainable_set_match_reference(H):
    z = _load_t5()
    model = H.VLT5(_t5_cfg(H))
    _load_state(model, z, torch.float32)
    assert sorted(H.trainable_names(model, model.config)) == sorted(str(n) for n in z["meta_trainable"])


def test_host_model_matches_reference_vlt5_cpu(H):
    """The reference's own VLT5 (tests/golden/make_golden_vlt5.py: T5-VL-PET-large flags, gate scale 0.3) against
    host.VLT5 with the eager T5 PET restatement on the sites: loss and every trainable gradient.  The model is fp64, but the
    reference's T5Attention takes its softmax in float32 (`F.softmax(scores.float())`, my_transformers/modeling_t5.py:655)
    and T5LayerNorm its variance (235-252), so the golden itself carries ~1e-8 of fp32 noise: bars 1e-7 / 2e-6."""
    from oracle.eager_ref import use_eager_pet
    z = _load_t5()
    model = use_eager_pet(H.VLT5(_t5_cfg(H)).double().eval())
    _load_state(model, z, torch.float64)
    names = [str(n) for n in z["meta_trainable"]]
    params = dict(model.named_parameters())
    for task in ("vqa", "nlvr"):
        model.zero_grad()
        loss = model.train_step(_batch(z, task, torch.float64))["loss"]
        loss.backward()
        assert abs(loss.item() - float(z[f"{task}/loss"])) < 1e-7, (loss.item(), float(z[f"{task}/loss"]))
        for n in names:
            assert rel(params[n].grad.numpy(), z[f"{task}/grad/{n}"]) < 2e-6, (task, n)


def test_t5_generate_matches_reference_cached_decode_cpu(H):
    """T5 twin of test_generate_matches_reference_cached_decode_cpu (golden: the reference VLT5's forward with
    past_key_values, tests/golden/make_golden_generate.py t5).  The reference takes softmax and RMS statistics in float32
    inside an fp64 model (see test_host_model_matches_reference_vlt5_cpu): logits bar 1e-6, tokens identical."""
    z, model, ids, vis, proc = _generate_case(H, torch.float64, "cpu", eager=True, arch="t5")
    tokens, logits = model.generate(ids, vis, task="vqa", max_length=int(z["meta_max_length"]), min_length=int(z["meta_min_length"]),
                                    logits_processor=proc, return_step_logits=True)
    assert tokens.tolist() == z["vqa/tokens"].tolist()
    assert rel(logits.numpy(), z["vqa/step_logits"]) < 1e-6
    with torch.no_grad():          # the cached steps reproduce the model's own full teacher-forced pass
        enc, mask = model.encoder(ids, vis, task="vqa")
        h = model.decoder(tokens[:, :-1], enc, encoder_mask=mask, task="vqa")
        full = model.lm_head(h * (model.model_dim ** -0.5))
    assert rel(full.numpy(), logits.numpy()) < 1e-9


@pytest.mark.gpu
def test_t5_generate_with_cuda_vpa_matches_reference_cached_decode(H):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z, model, ids, vis, proc = _generate_case(H, torch.float32, "cuda", eager=False, arch="t5")
    tokens, logits = model.generate(ids, vis, task="vqa", max_length=int(z["meta_max_length"]), min_length=int(z["meta_min_length"]),
                                    logits_processor=proc, return_step_logits=True)
    assert tokens.cpu().tolist() == z["vqa/tokens"].tolist()
    assert rel(logits.double().cpu().numpy(), z["vqa/step_logits"]) < 2e-4


def test_zero_init_flags_follow_the_reference_rules(H):
    """host.weight_initialization restates trainer_base.py:544-599: with the T5 scripts' zero-init flags the up projections of
    adapter, gate and decoder VPA start at zero, so the PET-augmented model computes the frozen backbone's function at step 0
    (large gate: y1 = x2 and G = sigmoid(0) = 1/2 scaled by s; the VPA adds nothing)."""
    from oracle.eager_ref import use_eager_pet
    cfg = _t5_cfg(H, use_encoder_multihead_up_zero_init=True, use_encoder_gating_large_x_lowrank_up_zero_init=True,
                  use_decoder_enc_vpa_up_zero_init=True)
    model = use_eager_pet(H.VLT5(cfg).double().eval())
    z = _load_t5()
    _load_state(model, z, torch.float64)
    done = H.weight_initialization(model, cfg)
    keys = [str(k) for k in z["meta_state_keys"]]
    want = [k for k in keys if "adapter_multihead_up" in k or "adapter_gating_large_x_up" in k or
            ("attn_value_parallel_adapter" in k and "up_sampler" in k)]
    params = dict(model.named_parameters())
    assert done and set(done) <= set(want) and all(float(params[n].detach().abs().sum()) == 0.0 for n in done)
    assert {params[n].data_ptr() for n in done} == {model.state_dict()[k].data_ptr() for k in want}     # aliased task keys included
    assert H.weight_initialization(model, _t5_cfg(H)) == []
    small = H.VLBart(_cfg(H, "small", use_encoder_gating_small_up_zero_init=True))
    zs = H.weight_initialization(small, small.config)
    assert zs and all("adapter_gating_small_xy_cat" in n for n in zs)


def test_test_step_handles_nlvr_pairs_cpu(H):
    """test_step on an NLVR batch (two images per sample, nlvr_model.py:156-176): both host models flatten the pair and add the
    image-order ids, as their train_step does."""
    from oracle.eager_ref import use_eager_pet
    for arch in ("bart", "t5"):
        z = _load("large") if arch == "bart" else _load_t5()
        model = use_eager_pet((H.VLBart(_cfg(H, "large")) if arch == "bart" else H.VLT5(_t5_cfg(H))).double().eval())
        _load_state(model, z, torch.float64)
        b = _batch(z, "nlvr", torch.float64)
        out = model.test_step({k: v for k, v in b.items() if k != "target_ids"}, max_length=5, min_length=5)["token_ids"]
        assert out.shape == (b["input_ids"].shape[0], 5)
        out5 = model.test_step({k: v for k, v in b.items() if k != "target_ids"}, max_length=5, min_length=5, num_beams=3)["token_ids"]
        assert out5.shape == out.shape


def test_t5_base_trainable_count_is_the_reference_checksum(H):
    """T5-base + VL-PET-large at r = rg = dec_r = 96 (BASELINE config 3): 10 499 712 trainable parameters, the count obtained
    by instantiating the reference VLT5."""
    cfg = H.t5_base_vlpet_large()
    with torch.device("meta"):
        model = H.VLT5(cfg)
    names = set(H.trainable_names(model, cfg))
    assert sum(p.numel() for k, p in model.named_parameters() if k in names) == 10499712


@pytest.mark.gpu
def test_host_t5_with_cuda_pet_matches_reference_vlt5(H):
    """T5 PET sites (K1 with s = 0.3, no LayerNorm behind), the T5 value parallel adapter (K2) and the RMS-norm visual
    projection (K3) through the CUDA kernels inside host.VLT5, against the reference VLT5 golden (fp32)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200 as V
    z = _load_t5()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = H.VLT5(_t5_cfg(H)).eval()
    _load_state(model, z, torch.float32)
    model.cuda()
    names = [str(n) for n in z["meta_trainable"]]
    params = dict(model.named_parameters())
    for n in names:
        params[n].requires_grad_(True)
    n0 = V.launch_count()
    for task in ("vqa", "nlvr"):
        model.zero_grad()
        loss = model.train_step(_batch(z, task, torch.float32))["loss"]
        loss.backward()
        assert abs(loss.item() - float(z[f"{task}/loss"])) < 5e-5 * abs(float(z[f"{task}/loss"]))
        for n in names:
            assert rel(params[n].grad.double().cpu().numpy(), z[f"{task}/grad/{n}"]) < 5e-4, (task, n)
    assert V.launch_count() - n0 >= 2 * (2 * 2 * 2 + 2 + 1), "PET sites did not run the CUDA kernels"


def test_beam_search_equals_transformers_generate_on_a_stock_model():
    """The search loop of host/generation.py against the installed transformers' own ``generate(num_beams=...)`` (the descendant of
    the 4.2.1 loop the reference pins) on a stock, randomly initialised tiny BART: same tokens for several beam widths, length
    penalties and both early-stopping modes, with hypotheses that end at different lengths.  A shared additive logits processor
    makes a random-init model emit varied tokens and EOS.  Library model on CPU: no PET kernels involved."""
    from transformers import BartConfig, BartForConditionalGeneration, LogitsProcessor, LogitsProcessorList
    import vlpet_b200.host.generation as G
    import sys
    if "ref_import" in sys.modules:              # an earlier test imported the reference: undo the init_weights shim for stock models
        sys.modules["ref_import"].remove_shims()
    torch.manual_seed(0)
    V, B, T = 30, 4, 9
    cfg = BartConfig(vocab_size=V, d_model=32, encoder_layers=1, decoder_layers=1, encoder_attention_heads=2,
                     decoder_attention_heads=2, encoder_ffn_dim=64, decoder_ffn_dim=64, max_position_embeddings=64, pad_token_id=1,
                     bos_token_id=0, eos_token_id=2, decoder_start_token_id=2, forced_eos_token_id=None, forced_bos_token_id=None)
    model = BartForConditionalGeneration(cfg).double().eval()
    lengths = set()
    for seed, nb, lp, es in ((1, 4, 1.0, False), (2, 4, 0.0, False), (3, 3, 1.0, True), (4, 5, 2.0, False), (5, 2, 0.5, True), (6, 1, 1.0, False)):
        g = torch.Generator().manual_seed(seed)
        ids = torch.randint(3, V, (B, 7), generator=g)
        bias = torch.randn(T, V, generator=g, dtype=torch.float64) * 2.0
        bias[:, 1] = -1e9                                        # never the pad token
        bias[:, 2] += 2.5                                        # EOS often among the candidates: hypotheses finish early

        class Bias(LogitsProcessor):
            def __call__(self, input_ids, scores):
                return scores + bias[input_ids.shape[1] - 1].to(scores.dtype)

        with torch.no_grad():
            want = model.generate(input_ids=ids, num_beams=nb, max_length=T, min_length=0, do_sample=False, early_stopping=es,
                                  length_penalty=lp, no_repeat_ngram_size=0, logits_processor=LogitsProcessorList([Bias()]))
            enc = model.get_encoder()(input_ids=ids).last_hidden_state.repeat_interleave(nb, 0)
        state = {"tok": None}

        def step(new):
            state["tok"] = new if state["tok"] is None else torch.cat([state["tok"], new], 1)
            return model(encoder_outputs=(enc,), decoder_input_ids=state["tok"]).logits[:, -1]

        def reorder(idx):
            state["tok"] = state["tok"].index_select(0, idx)

        got = G.beam_search(step, reorder, B, nb, "cpu", 2, 1, 2, T, 0, lp, es, logits_processor=lambda s, t, x: x + bias[s])
        w = max(want.shape[1], got.shape[1])
        pad = lambda x: torch.cat([x, torch.full((x.shape[0], w - x.shape[1]), 1)], 1)    # noqa: E731
        assert torch.equal(pad(want), pad(got)), (seed, nb, lp, es, want, got)
        lengths |= {int((row != 1).sum()) for row in got}
    assert len(lengths) >= 3, lengths                            # hypotheses of different lengths were compared
