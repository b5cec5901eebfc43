"""Model-parallel ILQL heads and model (reference: ``trlx/models/modeling_nemo_ilql.py`` — ``ParallelILQLHeads``
``:123-159``, ``LMHeads`` ``:162-212``, ``ILQLGPT`` ``:255-785``).

Built from :func:`trlx_b200.models.modeling_nemo_ppo.make_parallel_head`: the value head is column → row parallel (one
all-reduce), every Q head is column → column parallel with a vocabulary-wide, gathered output.  Target Q heads are
frozen copies updated by Polyak averaging on the local shards (no communication: both copies are sharded identically).
Activations are batch-first; the reference's ``[T, N, …] → [N, T, …]`` transposition is therefore a no-op here.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200.models.megatron_api import MegatronModelMixin, PipelineTensorAPI, unwrap_float16_module  # noqa: F401
from trlx_b200.models.modeling_nemo_ppo import ParallelLinear, make_parallel_head, reshard_for_pipeline_parallelism  # noqa: F401
from trlx_b200.parallel import state as parallel_state


class ParallelILQLHeads(nn.Module):
    """``(qs, target_qs, vs)`` from tensor-parallel MLP heads (reference ``:123-159``).  ``config`` is the
    :class:`~trlx_b200.models.modeling_ilql.ILQLConfig` (``two_qs``, ``alpha``)."""

    def __init__(self, config, hidden_size: int, vocab_size: int, sequence_parallel: bool = False,
                 dtype: torch.dtype = torch.bfloat16, tp: Optional[parallel_state.ModelParallelState] = None):
        super().__init__()
        self.config, self.hidden_size, self.vocab_size = config, hidden_size, vocab_size
        self.v_head = make_parallel_head(hidden_size, 1, sequence_parallel=sequence_parallel, dtype=dtype, tp=tp)
        n_qs = 2 if config.two_qs else 1
        self.q_heads = nn.ModuleList(make_parallel_head(hidden_size, vocab_size, dtype=dtype, tp=tp) for _ in range(n_qs))
        self.target_q_heads = nn.ModuleList(deepcopy(q) for q in self.q_heads)
        self.target_q_heads.requires_grad_(False)

    def forward(self, hidden_states: torch.Tensor) -> Tuple[Tuple[torch.Tensor, ...], Tuple[torch.Tensor, ...], torch.Tensor]:
        qs = tuple(q(hidden_states) for q in self.q_heads)
        with torch.no_grad():
            target_qs = tuple(q(hidden_states) for q in self.target_q_heads)
        return qs, target_qs, self.v_head(hidden_states)

    @torch.no_grad()
    def _sync_target_q_heads(self, alpha: float) -> None:
        for target, source in zip(self.target_q_heads, self.q_heads):
            for tp_, sp_ in zip(target.parameters(), source.parameters()):
                tp_.data.lerp_(sp_.data.to(tp_.dtype), alpha)  # alpha * q + (1 - alpha) * target

    def sync_target_q_heads(self) -> None:
        self._sync_target_q_heads(self.config.alpha)


class LMHeads(PipelineTensorAPI, nn.Module):
    """Language model + extra heads evaluated on its last hidden state (reference ``:162-212``): returns
    ``(logits, heads_output)``."""

    def __init__(self, language_model: nn.Module, other_heads: nn.Module):
        super().__init__()
        self.language_model, self.other_heads = language_model, other_heads

    def load_state_dict(self, lm_state_dict, strict: bool = True, **kw):
        """Load the language-model weights only (the heads are trained from scratch), as in the reference."""
        return self.language_model.load_state_dict(lm_state_dict, strict=strict, **kw)

    def forward(self, *args, **kwargs):
        out = self.language_model(*args, output_hidden_states=True, **kwargs)
        return out.logits, self.other_heads(out.hidden_states[-1])


class ILQLGPT(MegatronModelMixin, nn.Module):
    """Causal LM (tensor-parallel blocks) with :class:`ParallelILQLHeads`, plus the reference's advantage-shifted
    next-token distribution for generation (``:723-735``): ``log π_β + β · (min target-Q − V)``."""

    def __init__(self, ilql_config, config=None, language_model: Optional[nn.Module] = None, hidden_size: Optional[int] = None,
                 vocab_size: Optional[int] = None, dtype: torch.dtype = torch.bfloat16, metric_fn=None):
        super().__init__()
        self.ilql_config, self.config, self.metric_fn = ilql_config, config, metric_fn
        if language_model is None:
            from trlx_b200.models.modeling_base import build_base_model
            from trlx_b200.parallel.tensor_parallel import apply_tensor_parallel

            language_model = build_base_model(config.model.model_path, "causal", dtype=dtype)
            st = parallel_state.get_model_parallel()
            if st.tp_group is not None and st.tp_size > 1:
                apply_tensor_parallel(language_model, st.tp_group, st.tp_rank, st.tp_size,
                                      sequence_parallel=bool(getattr(config.train.parallel, "sequence_parallel", False)))
        cfg = getattr(language_model, "config", None)
        hidden_size = hidden_size or getattr(cfg, "hidden_size", None) or getattr(cfg, "n_embd")
        vocab_size = vocab_size or getattr(cfg, "vocab_size")
        self.model = LMHeads(language_model, ParallelILQLHeads(ilql_config, hidden_size, vocab_size, dtype=dtype))

    @property
    def heads(self) -> ParallelILQLHeads:
        return self.model.other_heads

    def forward(self, input_ids, attention_mask=None, position_ids=None, **kw):
        """``(logits, (qs, target_qs, vs))``"""
        return self.model(input_ids, attention_mask=attention_mask, position_ids=position_ids, **kw)

    @torch.no_grad()
    def shifted_logits(self, input_ids, attention_mask=None, position_ids=None, beta: Optional[float] = None) -> torch.Tensor:
        """Next-token scores used for sampling: ``log_softmax(logits) + β · (min_i target_q_i − V)`` at the last position."""
        logits, (_, target_qs, vs) = self.forward(input_ids, attention_mask=attention_mask, position_ids=position_ids)
        target_q = target_qs[0] if len(target_qs) == 1 else torch.minimum(target_qs[0], target_qs[1])
        beta = self.ilql_config.gen_kwargs.get("beta", 1.0) if beta is None else beta
        adv = target_q[:, -1, :].float() - vs[:, -1, :].float()
        return F.log_softmax(logits[:, -1, :].float(), -1) + beta * adv

    def sync_target_q_heads(self) -> None:
        self.heads.sync_target_q_heads()

    def _loss(self, batch):
        """ILQL loss of one :class:`~trlx_b200.data.ilql_types.ILQLBatch` (reference loss closure ``:612-683``): heads are
        evaluated on the state / action positions only."""
        from trlx_b200.models.modeling_ilql import batched_index_select

        dev = next(self.parameters()).device
        batch = type(batch)(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(batch).items()})
        out = self.model.language_model(batch.input_ids, attention_mask=batch.attention_mask, output_hidden_states=True)
        hs = out.hidden_states[-1]
        qs, target_qs, _ = self.heads(batched_index_select(hs, batch.actions_ixs, 1))
        vs = self.heads.v_head(batched_index_select(hs, batch.states_ixs, 1))
        return self.ilql_config.loss((out.logits, (qs, target_qs, vs)), batch)

    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, max_new_tokens: int = 16, beta: Optional[float] = None,
                 temperature: float = 1.0, do_sample: bool = False, eos_token_id: Optional[int] = None,
                 pad_token_id: int = 0, **_):
        """Sampling from the advantage-shifted distribution, one token at a time (reference ``:737-785``)."""
        ids = input_ids
        am = attention_mask if attention_mask is not None else torch.ones_like(ids)
        done = torch.zeros(ids.shape[0], dtype=torch.bool, device=ids.device)
        for _step in range(int(max_new_tokens)):
            pos = (am.long().cumsum(-1) - 1).clamp_min(0)
            scores = self.shifted_logits(ids, am, pos, beta) / max(float(temperature), 1e-6)
            nxt = torch.multinomial(torch.softmax(scores, -1), 1).squeeze(-1) if do_sample else scores.argmax(-1)
            nxt = torch.where(done, torch.full_like(nxt, pad_token_id), nxt)
            ids = torch.cat([ids, nxt.unsqueeze(-1)], 1)
            am = torch.cat([am, (~done).long().unsqueeze(-1)], 1)
            if eos_token_id is not None:
                done = done | (nxt == eos_token_id)
                if bool(done.all()):
                    break
        return ids
