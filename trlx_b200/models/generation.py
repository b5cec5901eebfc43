"""Sampling loop shared by every model wrapper (the PyTorch path).

Covers the ``generate`` keyword surface the reference passes to HF (``trlx/trainer/accelerate_ppo_trainer.py:89-101``,
``accelerate_base_trainer.py:256-282``): ``max_new_tokens`` / ``max_length`` / ``min_new_tokens`` / ``min_length``,
``do_sample``, ``temperature``, ``top_k``, ``top_p``, ``eos_token_id``, ``pad_token_id``, ``synced_gpus`` (accepted,
irrelevant here) for decoder-only and encoder-decoder models, with a KV cache.  On B200 the trainers route rollouts
through :mod:`trlx_b200.engine` (CUDA-graph decode on the sm_100a kernels) instead; this loop is the CPU path, the
oracle for the engine tests, and the fallback for options the engine does not implement.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Union

import torch
import torch.nn.functional as F


def top_k_top_p_filter(logits: torch.Tensor, top_k: int = 0, top_p: float = 1.0) -> torch.Tensor:
    """Mask (to -inf) everything outside the top-k tokens and outside the smallest nucleus with mass ≥ top_p."""
    if top_k and top_k > 0:
        k = min(int(top_k), logits.shape[-1])
        kth = torch.topk(logits, k, dim=-1).values[..., -1:]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and 0.0 < top_p < 1.0:
        sorted_logits, idx = torch.sort(logits, dim=-1, descending=True)
        cum = torch.softmax(sorted_logits, dim=-1).cumsum(-1)
        remove = cum - torch.softmax(sorted_logits, dim=-1) >= top_p  # keep the token that crosses the threshold
        remove[..., 0] = False
        logits = logits.masked_fill(torch.zeros_like(remove).scatter(-1, idx, remove), float("-inf"))
    return logits


def sample_next(logits: torch.Tensor, do_sample: bool, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
                generator: Optional[torch.Generator] = None) -> torch.Tensor:
    if not do_sample or temperature == 0:
        return logits.argmax(-1)
    logits = logits.float()
    if temperature is not None and temperature != 1.0:
        logits = logits / temperature
    logits = top_k_top_p_filter(logits, top_k, top_p)
    return torch.multinomial(torch.softmax(logits, -1), 1, generator=generator).squeeze(-1)


def sample_sync_groups(model) -> list:
    """``[(group, src)]`` of the model-parallel groups that must agree on sampled tokens (tensor / pipeline parallel
    replicas of one model); empty for plain data parallelism."""
    for m in model.modules():
        groups = getattr(m, "_sample_sync", None)
        if groups:
            return list(groups)
    return []


def sync_tokens(tokens: torch.Tensor, groups: list) -> torch.Tensor:
    if groups:
        import torch.distributed as dist

        tokens = tokens.contiguous()
        for group, src in groups:
            dist.broadcast(tokens, src=src, group=group)
    return tokens


def _ids(x) -> List[int]:
    if x is None:
        return []
    if isinstance(x, (list, tuple)):
        return [int(i) for i in x]
    return [int(x)]


def _reorder_cache(past, idx: torch.Tensor):
    """Gather the batch dimension of every tensor in a (nested) KV-cache structure."""
    if past is None:
        return None
    if torch.is_tensor(past):
        return past.index_select(0, idx)
    return type(past)(_reorder_cache(p, idx) for p in past)


@torch.no_grad()
def beam_search(model, input_ids, attention_mask=None, num_beams: int = 4, max_new_tokens=None, max_length=None,
                min_new_tokens: int = 0, min_length: int = 0, eos_token_id=None, pad_token_id=None, logits_processor=None,
                decoder_start_token_id=None, length_penalty: float = 1.0, early_stopping: bool = False,
                temperature: float = 1.0) -> torch.Tensor:
    """Deterministic beam search (``num_beams > 1``, the ``gen_experience_kwargs`` of the reference's translation example,
    ``examples/ppo_translation_t5.py:78-83``) for decoder-only and encoder-decoder models with a KV cache.

    Hypotheses live in a flat ``[B·k]`` batch; a finished beam is frozen (it may only emit padding at zero cost) so it keeps
    competing with live beams by score, and the final ranking divides the summed log-probability by
    ``length ** length_penalty`` like HF's ``BeamSearchScorer``."""
    cfg = model.config
    eos_ids = _ids(eos_token_id if eos_token_id is not None else getattr(cfg, "eos_token_id", None))
    if pad_token_id is None:
        pad_token_id = getattr(cfg, "pad_token_id", None)
    if pad_token_id is None:
        pad_token_id = eos_ids[0] if eos_ids else 0
    device = input_ids.device
    B, Q = input_ids.shape
    k = num_beams
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    seq2seq = bool(getattr(cfg, "is_encoder_decoder", False))
    if max_new_tokens is None:
        max_new_tokens = 20 if max_length is None else max(max_length - (1 if seq2seq else Q), 0)
    min_new = max(int(min_new_tokens or 0), int(min_length or 0) - (1 if seq2seq else Q), 0)
    eos_t = torch.tensor(eos_ids, device=device, dtype=torch.long) if eos_ids else None
    sync = sample_sync_groups(model)

    expand = torch.arange(B, device=device).repeat_interleave(k)
    mask = attention_mask.long().index_select(0, expand)
    if seq2seq:
        start = decoder_start_token_id if decoder_start_token_id is not None else getattr(cfg, "decoder_start_token_id", None)
        start = pad_token_id if start is None else start
        enc = model.encode(input_ids=input_ids, attention_mask=attention_mask)
        enc = _reorder_cache(enc, expand) if not torch.is_tensor(enc) else enc.index_select(0, expand)
        seqs = torch.full((B * k, 1), int(start), dtype=torch.long, device=device)
        step_in = seqs
    else:
        seqs = input_ids.index_select(0, expand)
        positions = (mask.cumsum(-1) - 1).clamp_min(0)
        step_in, step_pos = seqs, positions
    scores = torch.zeros(B, k, device=device)
    scores[:, 1:] = float("-inf")  # all beams of a row start identical: only the first may spawn children
    finished = torch.zeros(B * k, dtype=torch.bool, device=device)
    lengths = torch.zeros(B * k, dtype=torch.long, device=device)
    past = None
    for step in range(max_new_tokens):
        if seq2seq:
            out = model.decode(decoder_input_ids=step_in, encoder_hidden_states=enc, attention_mask=mask, past_key_values=past,
                               use_cache=True)
        else:
            out = model(input_ids=step_in, attention_mask=mask, position_ids=step_pos, past_key_values=past, use_cache=True)
        past = out.past_key_values
        if not seq2seq and past is not None and past[0][0].shape[2] > mask.shape[1]:
            extra = past[0][0].shape[2] - mask.shape[1]  # virtual tokens of prompt/prefix adapters
            mask = torch.cat([torch.ones(B * k, extra, dtype=mask.dtype, device=device), mask], 1)
        logits = out.logits[:, -1, :].float()
        if temperature not in (None, 1.0, 0, 0.0):
            logits = logits / temperature
        if step < min_new and eos_ids:
            logits[:, eos_ids] = float("-inf")
        for proc in logits_processor or []:
            logits = proc(seqs, logits)
        logp = torch.log_softmax(logits, -1)
        V = logp.shape[-1]
        # a finished hypothesis can only continue with padding, at no cost
        frozen = torch.full_like(logp, float("-inf"))
        frozen[:, pad_token_id] = 0.0
        logp = torch.where(finished[:, None], frozen, logp)
        cand = (scores.view(B * k, 1) + logp).view(B, k * V)
        top_scores, top_idx = cand.topk(k, dim=-1)
        beam_src = top_idx // V  # [B, k] index of the parent beam
        token = (top_idx % V).view(-1)
        parent = (beam_src + torch.arange(B, device=device)[:, None] * k).view(-1)
        token = sync_tokens(token, sync)
        parent = sync_tokens(parent, sync)
        scores = top_scores
        was_finished = finished.index_select(0, parent)
        token = torch.where(was_finished, torch.full_like(token, pad_token_id), token)
        seqs = torch.cat([seqs.index_select(0, parent), token[:, None]], 1)
        lengths = lengths.index_select(0, parent) + (~was_finished).long()
        finished = was_finished | (torch.isin(token, eos_t) if eos_t is not None else torch.zeros_like(was_finished))
        past = _reorder_cache(past, parent)
        if seq2seq:
            step_in = token[:, None]
        else:
            mask = torch.cat([mask.index_select(0, parent), torch.ones(B * k, 1, dtype=mask.dtype, device=device)], 1)
            step_pos = step_pos.index_select(0, parent)[:, -1:] + 1
            step_in = token[:, None]
        if early_stopping and bool(finished.view(B, k).any(-1).all()):
            break
        if bool(finished.all()):
            break
    norm = scores / lengths.clamp_min(1).view(B, k).float().pow(length_penalty)
    best = norm.argmax(-1) + torch.arange(B, device=device) * k
    return seqs.index_select(0, best)


@torch.no_grad()
def generate(model, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
             max_new_tokens: Optional[int] = None, max_length: Optional[int] = None, min_new_tokens: int = 0,
             min_length: int = 0, do_sample: bool = False, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
             eos_token_id: Union[int, Sequence[int], None] = None, pad_token_id: Optional[int] = None,
             logits_processor: Optional[List[Callable]] = None, decoder_start_token_id: Optional[int] = None,
             generator: Optional[torch.Generator] = None, num_beams: int = 1, synced_gpus: bool = False,
             inputs: Optional[torch.Tensor] = None, **unused) -> torch.Tensor:
    """Returns ``[B, prompt_len + n_new]`` token ids (decoder-only) or the decoder sequence incl. its start token
    (encoder-decoder).  Finished rows are filled with ``pad_token_id``."""
    if input_ids is None:
        input_ids = inputs
    if num_beams and num_beams > 1:
        return beam_search(model, input_ids, attention_mask, num_beams=int(num_beams), max_new_tokens=max_new_tokens,
                           max_length=max_length, min_new_tokens=min_new_tokens, min_length=min_length,
                           eos_token_id=eos_token_id, pad_token_id=pad_token_id, logits_processor=logits_processor,
                           decoder_start_token_id=decoder_start_token_id,
                           length_penalty=float(unused.get("length_penalty", 1.0)),
                           early_stopping=bool(unused.get("early_stopping", False)), temperature=temperature)
    cfg = model.config
    eos_ids = _ids(eos_token_id if eos_token_id is not None else getattr(cfg, "eos_token_id", None))
    if pad_token_id is None:
        pad_token_id = getattr(cfg, "pad_token_id", None)
    if pad_token_id is None:
        pad_token_id = eos_ids[0] if eos_ids else 0
    device = input_ids.device
    B, Q = input_ids.shape
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    seq2seq = bool(getattr(cfg, "is_encoder_decoder", False))

    if max_new_tokens is None:
        if max_length is None:
            max_new_tokens = 20
        else:
            max_new_tokens = max(max_length - (1 if seq2seq else Q), 0)
    min_new = max(int(min_new_tokens or 0), int(min_length or 0) - (1 if seq2seq else Q), 0)

    finished = torch.zeros(B, dtype=torch.bool, device=device)
    sync = sample_sync_groups(model)
    eos_t = torch.tensor(eos_ids, device=device, dtype=torch.long) if eos_ids else None

    if seq2seq:
        start = decoder_start_token_id if decoder_start_token_id is not None else getattr(cfg, "decoder_start_token_id", None)
        if start is None:
            start = pad_token_id
        enc = model.encode(input_ids=input_ids, attention_mask=attention_mask)
        seqs = torch.full((B, 1), int(start), dtype=torch.long, device=device)
        past = None
        step_in = seqs
        for step in range(max_new_tokens):
            out = model.decode(decoder_input_ids=step_in, encoder_hidden_states=enc, attention_mask=attention_mask,
                               past_key_values=past, use_cache=True)
            past = out.past_key_values
            logits = out.logits[:, -1, :].float()
            if step < min_new and eos_ids:
                logits[:, eos_ids] = float("-inf")
            for proc in logits_processor or []:
                logits = proc(seqs, logits)
            nxt = sync_tokens(sample_next(logits, do_sample, temperature, top_k, top_p, generator), sync)
            nxt = torch.where(finished, torch.full_like(nxt, pad_token_id), nxt)
            seqs = torch.cat([seqs, nxt[:, None]], 1)
            if eos_t is not None:
                finished = finished | torch.isin(nxt, eos_t)
            step_in = nxt[:, None]
            if bool(finished.all()):
                break
        return seqs

    mask = attention_mask.long()
    positions = (mask.cumsum(-1) - 1).clamp_min(0)
    seqs = input_ids
    past = None
    step_ids, step_pos = input_ids, positions
    for step in range(max_new_tokens):
        out = model(input_ids=step_ids, attention_mask=mask, position_ids=step_pos, past_key_values=past, use_cache=True)
        past = out.past_key_values
        if past is not None and past[0][0].shape[2] > mask.shape[1]:
            # adapters (prefix / prompt tuning) put extra virtual keys in the cache on the first call
            extra = past[0][0].shape[2] - mask.shape[1]
            mask = torch.cat([torch.ones(B, extra, dtype=mask.dtype, device=device), mask], 1)
        logits = out.logits[:, -1, :].float()
        if step < min_new and eos_ids:
            logits[:, eos_ids] = float("-inf")
        for proc in logits_processor or []:
            logits = proc(seqs, logits)
        nxt = sync_tokens(sample_next(logits, do_sample, temperature, top_k, top_p, generator), sync)
        nxt = torch.where(finished, torch.full_like(nxt, pad_token_id), nxt)
        seqs = torch.cat([seqs, nxt[:, None]], 1)
        if eos_t is not None:
            finished = finished | torch.isin(nxt, eos_t)
        mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype, device=device)], 1)
        step_pos = step_pos[:, -1:] + 1
        step_ids = nxt[:, None]
        if bool(finished.all()):
            break
    return seqs
