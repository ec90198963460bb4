"""Golden fixture for SURVEY section 8 row f-4: the reference's KV-cached decode path -- ``VLBart.forward`` with
``past_key_values`` / ``use_cache`` (src/modeling_bart.py:1522-1602, my_transformers/modeling_bart.py:1617-1788), where the
decoder's cross-attention keys / values, the values already carrying the value parallel adapter
(my_transformers/modeling_bart.py:419-430), are computed at the first step and reused from the cache afterwards -- driven by
a plain greedy loop (what ``test_step`` -> ``generate(num_beams=1)`` does for VQA / GQA / NLVR, multitask.py:480, 516;
vqa_model.py:235-288).  Test infrastructure; run in the build container:

    python tests/golden/make_golden_generate.py

Writes ``vlbart_tiny_generate.npz`` and ``vlt5_tiny_generate.npz`` (the T5 twin: src/modeling_t5.py:560-690 with
``past_key_values``, my_transformers/modeling_t5.py:588-613 for the cached value parallel adapter): state_dict, batch, the greedy
token ids and the last-position logits of every step.
HF's search machinery (logits processors, beam search) is the caller's and is not part of the fixture.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_vlbart as MV  # noqa: E402

MAX_LEN = 8
MIN_LEN = 6      # HF MinLengthLogitsProcessor: the EOS logit is -inf while the sequence is shorter (a random-init model emits EOS at once)


@torch.no_grad()
def greedy(model, config, ids, vis_inputs, task, bias):
    B = ids.shape[0]
    encoder = model.model.encoder if hasattr(model, "model") else model.encoder
    enc = encoder(input_ids=ids, vis_inputs=vis_inputs, return_dict=True, task=task)
    amask = ids.ne(config.pad_token_id).to(torch.float64)      # what HF generate passes along with encoder_outputs
    dec = torch.full((B, 1), config.decoder_start_token_id, dtype=torch.long)
    past, steps = None, []
    done = torch.zeros(B, dtype=torch.bool)
    while dec.shape[1] < MAX_LEN:
        o = model(input_ids=None, attention_mask=amask, vis_inputs=vis_inputs, encoder_outputs=enc,
                  decoder_input_ids=dec if past is None else dec[:, -1:], past_key_values=past, use_cache=True, return_dict=True, task=task)
        logits = o.logits[:, -1]
        steps.append(logits.numpy().copy())
        past = o.past_key_values
        scores = logits + bias[:, len(steps) - 1]        # a logits processor (HF generate's hook): makes the tokens of a random-init model vary
        if dec.shape[1] < MIN_LEN:
            scores[:, config.eos_token_id] = -float("inf")
        nxt = scores.argmax(-1)
        nxt = torch.where(done, torch.full_like(nxt, config.pad_token_id), nxt)
        done |= nxt == config.eos_token_id
        dec = torch.cat([dec, nxt[:, None]], 1)
        if bool(done.all()):
            break
    # the same tokens without a cache: one full teacher-forced pass must reproduce every step's logits
    full = model(input_ids=None, attention_mask=amask, vis_inputs=vis_inputs, encoder_outputs=enc, decoder_input_ids=dec[:, :-1],
                 use_cache=False, return_dict=True, task=task).logits
    for t, s in enumerate(steps):
        assert np.abs(full[:, t].numpy() - s).max() < 1e-9, "reference cache path differs from its own full pass"
    return dec.numpy(), np.stack(steps, 1)


def main_t5():
    import make_golden_vlt5 as MT
    model, config = MT.build()
    model = model.double().eval()
    g = torch.Generator().manual_seed(6)
    B, Lt = 3, 7
    out = {}
    sd = model.state_dict()
    out["meta_state_keys"] = np.array(list(sd.keys()))
    for k, v in sd.items():
        out["sd/" + k] = v.detach().cpu().numpy()
    ids = torch.randint(3, 300, (B, Lt), generator=g)
    feats = torch.randn(B, 49, MT.FEAT, generator=g, dtype=torch.float64)
    boxes = torch.zeros(B, 49, 4, dtype=torch.float64)
    bias = 4.0 * torch.randn(B, MAX_LEN, 300, generator=g, dtype=torch.float64)     # T5 logits are O(1)
    tokens, logits = greedy(model, config, ids, (feats, boxes), "vqa", bias)
    out["vqa/input_ids"], out["vqa/vis_feats"], out["vqa/boxes"] = ids.numpy(), feats.numpy(), boxes.numpy()
    out["vqa/tokens"], out["vqa/step_logits"], out["vqa/logit_bias"] = tokens, logits, bias.numpy()
    out["meta_max_length"], out["meta_min_length"] = np.array(MAX_LEN), np.array(MIN_LEN)
    path = os.path.join(HERE, "vlt5_tiny_generate.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", "tokens", tokens.tolist(), "steps", logits.shape)


def main():
    model, config = MV.build("large")
    g = torch.Generator().manual_seed(5)
    B, Lt = 3, 7
    out = {}
    sd = model.state_dict()
    out["meta_state_keys"] = np.array(list(sd.keys()))
    for k, v in sd.items():
        out["sd/" + k] = v.detach().cpu().numpy()
    ids = torch.randint(3, 300, (B, Lt), generator=g)
    feats = torch.randn(B, 49, MV.FEAT, generator=g, dtype=torch.float64)
    boxes = torch.zeros(B, 49, 4, dtype=torch.float64)
    bias = 0.5 * torch.randn(B, MAX_LEN, 300, generator=g, dtype=torch.float64)
    tokens, logits = greedy(model, config, ids, (feats, boxes), "vqa", bias)
    out["vqa/logit_bias"] = bias.numpy()
    out["vqa/input_ids"], out["vqa/vis_feats"], out["vqa/boxes"] = ids.numpy(), feats.numpy(), boxes.numpy()
    out["vqa/tokens"], out["vqa/step_logits"] = tokens, logits
    out["meta_max_length"], out["meta_min_length"] = np.array(MAX_LEN), np.array(MIN_LEN)
    path = os.path.join(HERE, "vlbart_tiny_generate.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", "tokens", tokens.tolist(), "steps", logits.shape)


if __name__ == "__main__":
    if "t5" in sys.argv[1:]:
        main_t5()          # (separate processes: the BART and T5 import shims patch transformers differently)
    else:
        main()
