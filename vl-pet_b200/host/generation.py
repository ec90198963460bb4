"""Beam search over the KV-cached decode step of the host models (the caption task's
``generate(num_beams=5, max_length=40)``, multitask.py:587-588, caption_model.py test_step).

The search loop restates the algorithm of the reference's pinned ``transformers`` 4.2.1 (``GenerationMixin.beam_search`` +
``BeamSearchScorer`` / ``BeamHypotheses``): log-softmax scores, ``2 * num_beams`` candidates per sample and step, finished
hypotheses ranked by ``sum_logprobs / len ** length_penalty``, a sample is done when its worst kept hypothesis beats the best
score still reachable (or, with ``early_stopping``, as soon as ``num_beams`` hypotheses finished).  That library is not
importable offline in this image (transformers 5.5 removed the scorer), so the tests pin the loop three ways: token-for-token
against the installed transformers' own ``generate(num_beams=...)`` on a stock tiny BART (beam widths, length penalties, both
early-stopping modes, hypotheses of different lengths), by exhaustive enumeration on a restricted vocabulary, and by re-scoring
its output with a full teacher-forced pass (which also checks the beam re-ordering of the self-attention caches).  Host-side plumbing: PyTorch, no kernels of its own; the PET work inside a step is the model's.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch


class BeamHypotheses:
    """The ``num_beams`` best finished hypotheses of one sample (transformers 4.2.1 generation_beam_search.py:332-377)."""

    def __init__(self, num_beams: int, length_penalty: float, early_stopping: bool):
        self.num_beams, self.length_penalty, self.early_stopping = num_beams, length_penalty, early_stopping
        self.beams: List[Tuple[float, torch.Tensor, float]] = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp: torch.Tensor, sum_logprobs: float):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp, sum_logprobs))
            if len(self) > self.num_beams:
                order = sorted((s, i) for i, (s, _, _) in enumerate(self.beams))
                del self.beams[order[0][1]]
                self.worst_score = order[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


@torch.no_grad()
def beam_search(step: Callable[[torch.Tensor], torch.Tensor], reorder: Callable[[torch.Tensor], None], batch: int, num_beams: int,
                device, start_token: int, pad_token: int, eos_token: int, max_length: int, min_length: int = 0,
                length_penalty: float = 1.0, early_stopping: bool = False, logits_processor=None, return_scores: bool = False):
    """``step(new_tokens [batch * num_beams, T]) -> logits of the last position [batch * num_beams, V]`` feeds the cache;
    ``reorder(beam_idx [batch * num_beams])`` re-orders every cached tensor along the batch axis.  Returns the best
    hypothesis per sample, right-padded ([batch, <= max_length]) and, optionally, its sum of log-probabilities."""
    nb = num_beams
    tokens = torch.full((batch * nb, 1), start_token, dtype=torch.long, device=device)
    beam_scores = torch.zeros(batch, nb, dtype=torch.float64, device=device)
    beam_scores[:, 1:] = -1e9                                   # all beams start identical: only the first one may expand
    beam_scores = beam_scores.view(-1)
    hyps = [BeamHypotheses(nb, length_penalty, early_stopping) for _ in range(batch)]
    done = [False] * batch
    first = True
    while tokens.shape[1] < max_length:
        cur_len = tokens.shape[1]
        logits = step(tokens if first else tokens[:, -1:])
        first = False
        logp = torch.log_softmax(logits.double(), dim=-1)
        if logits_processor is not None:
            logp = logits_processor(cur_len - 1, tokens, logp)
        if cur_len < min_length:
            logp[:, eos_token] = -float("inf")
        V = logp.shape[-1]
        scores = (logp + beam_scores[:, None]).view(batch, nb * V)
        top_scores, top_idx = torch.topk(scores, 2 * nb, dim=1, largest=True, sorted=True)
        top_beam, top_tok = (top_idx // V).tolist(), (top_idx % V).tolist()
        top_scores_l = top_scores.tolist()
        next_scores = torch.zeros(batch, nb, dtype=torch.float64)
        next_tokens = torch.full((batch, nb), pad_token, dtype=torch.long)
        next_beams = torch.zeros(batch, nb, dtype=torch.long)
        for b in range(batch):
            if done[b]:
                next_beams[b] = b * nb                          # padding beams of a finished sample
                continue
            k = 0
            for rank in range(2 * nb):
                tok, src, sc = top_tok[b][rank], b * nb + top_beam[b][rank], top_scores_l[b][rank]
                if tok == eos_token:
                    if rank < nb:                               # an EOS outside the first num_beams candidates is dropped
                        hyps[b].add(tokens[src].clone(), sc)
                else:
                    next_scores[b, k], next_tokens[b, k], next_beams[b, k] = sc, tok, src
                    k += 1
                if k == nb:
                    break
            done[b] = done[b] or hyps[b].is_done(top_scores_l[b][0], cur_len)
        beam_idx = next_beams.view(-1).to(device)
        beam_scores = next_scores.view(-1).to(device)
        tokens = torch.cat([tokens.index_select(0, beam_idx), next_tokens.view(-1, 1).to(device)], dim=1)
        reorder(beam_idx)
        if all(done):
            break
    for b in range(batch):                                      # finalize: open beams become hypotheses
        if done[b]:
            continue
        for k in range(nb):
            hyps[b].add(tokens[b * nb + k].clone(), float(beam_scores[b * nb + k]))
    best = [max(h.beams, key=lambda x: x[0]) for h in hyps]
    out_len = min(max(t.shape[-1] for _, t, _ in best) + 1, max_length)
    out = torch.full((batch, out_len), pad_token, dtype=torch.long, device=device)
    for b, (_, t, _) in enumerate(best):
        out[b, :t.shape[-1]] = t
        if t.shape[-1] < max_length:
            out[b, t.shape[-1]] = eos_token                     # HF appends EOS to a hypothesis that ended before max_length
    if return_scores:
        return out, torch.tensor([sl for _, _, sl in best], dtype=torch.float64)
    return out


class NoRepeatNGram:
    """HF's NoRepeatNGramLogitsProcessor (``no_repeat_ngram_size``; 3 in facebook/bart-base's generation defaults, which the
    reference's ``generate`` calls inherit): a token that would complete an n-gram already present in the sequence gets
    -inf.  Usable as ``logits_processor=NoRepeatNGram(3)`` (or chained through ``chain``) in ``generate``."""

    def __init__(self, n: int):
        if n < 1:
            raise ValueError("no_repeat_ngram_size must be >= 1")
        self.n = n

    def __call__(self, step: int, tokens: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        n, cur_len = self.n, tokens.shape[1]
        if cur_len + 1 < n:
            return scores
        rows = tokens.tolist()
        for i, seq in enumerate(rows):
            prefix = tuple(seq[cur_len - n + 1:]) if n > 1 else ()
            banned = [seq[j + n - 1] for j in range(cur_len - n + 1) if tuple(seq[j:j + n - 1]) == prefix]
            if banned:
                scores[i, banned] = -float("inf")
        return scores


def chain(*processors):
    """Apply several ``logits_processor`` callables in order."""
    def run(step, tokens, scores):
        for p in processors:
            scores = p(step, tokens, scores)
        return scores
    return run
