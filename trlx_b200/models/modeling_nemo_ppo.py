"""Model-parallel building blocks of the PPO policy (reference: ``trlx/models/modeling_nemo_ppo.py``).

The reference assembles its tensor/pipeline-parallel PPO model from Apex / NeMo parts (``ParallelLinear`` ``:67-127``,
``make_parallel_head`` ``:130-149``, ``ValueHead`` ``:152-164``, ``RefLMHeads`` ``:167-312``,
``reshard_for_pipeline_parallelism`` ``:321-352``, ``PPOGPT`` ``:355-1222``).  This module provides the same pieces on this
framework's own parallel layer:

* the layout comes from :mod:`trlx_b200.parallel.state` (published by the ``Runtime``), collectives are
  ``torch.distributed`` calls wrapped in autograd functions, and on CUDA the GEMMs inside are the tcgen05 kernels behind
  ``ops.linear``;
* weights are initialised as slices of the *dense* initialisation (same global RNG stream on every rank), so a head
  built at ``tp = 4`` is numerically the ``tp = 1`` head, sharded — checkpoints convert by plain slicing / concatenation
  (:func:`shard_head_state_dict`);
* activations are batch-first ``[B, T, H]`` (the reference's Megatron layout is ``[T, B, H]``; pass ``seq_first=True`` to
  :class:`ValueHead` for that layout);
* :class:`RefLMHeads` keeps the reference policy as pinned host copies of the trainable parameters and swaps them in for
  the reference forward — the alternative to the hydra branch when *every* layer is trained and a second full model
  should not live in HBM.
"""
from __future__ import annotations

import contextlib
import re
from math import sqrt
from typing import Any, Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from trlx_b200.models.megatron_api import (MegatronBatchSampler, MegatronModelMixin,  # noqa: F401  (re-exported)
                                           PipelineTensorAPI, patch_attention_for_llama, unwrap_float16_module)
from trlx_b200.parallel import state as parallel_state
from trlx_b200.parallel.tensor_parallel import _CopyToTP, _ReduceFromTP
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


# ---- last-dimension collectives ----------------------------------------------------------------------------------------------
class _GatherLastDim(torch.autograd.Function):
    """all-gather along the feature dimension forward, keep the local slice backward (``gather_output=True``)."""

    @staticmethod
    def forward(ctx, x, group, rank):
        ctx.group, ctx.rank = group, rank
        world = dist.get_world_size(group)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=-1)

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        return g.chunk(world, dim=-1)[ctx.rank].contiguous(), None, None


class _SplitLastDim(torch.autograd.Function):
    """keep the local slice of the feature dimension forward, all-gather backward (row-parallel layer fed a full input)."""

    @staticmethod
    def forward(ctx, x, group, rank):
        ctx.group, ctx.rank = group, rank
        return x.chunk(dist.get_world_size(group), dim=-1)[rank].contiguous()

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g.contiguous(), group=ctx.group)
        return torch.cat(parts, dim=-1), None, None


def _linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    if x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16:
        from trlx_b200 import ops

        return ops.linear(x, w, b)
    return F.linear(x, w, b)


class ParallelLinear(nn.Module):
    """Linear layer sharded over its longer dimension (reference ``:67-127``): **column-parallel** (output features split,
    ``[out/tp, in]``) when ``in_size < out_size``, otherwise **row-parallel** (input features split, ``[out, in/tp]``, partial
    products all-reduced, bias added once after the reduction).

    ``gather_output`` (column): return the full output instead of this rank's slice.  ``input_is_parallel`` (row): the input
    is already the local feature slice (the output of a non-gathered column-parallel layer).  With a tensor-parallel size of
    one the layer is a plain ``nn.Linear`` with the same initialisation."""

    def __init__(self, in_size: int, out_size: int, init_method: Optional[Callable] = None, use_cpu_initialization: bool = False,
                 bias: bool = True, sequence_parallel: bool = False, gradient_accumulation_fusion: bool = False,
                 gather_output: bool = True, input_is_parallel: bool = False, dtype: torch.dtype = torch.bfloat16,
                 tp: Optional[parallel_state.ModelParallelState] = None):
        super().__init__()
        st = tp or parallel_state.get_model_parallel()
        self.group, self.rank, self.world = st.tp_group, st.tp_rank, (st.tp_size if st.tp_group is not None else 1)
        self.in_size, self.out_size = in_size, out_size
        self.column = in_size < out_size
        self.gather_output, self.input_is_parallel = gather_output, input_is_parallel
        self.sequence_parallel = sequence_parallel
        split = out_size if self.column else in_size
        if split % self.world:
            raise ValueError(f"ParallelLinear: {split} features are not divisible by the tensor-parallel size {self.world}")
        # dense initialisation on every rank (identical RNG stream), then keep the local slice
        full_w = torch.empty(out_size, in_size, dtype=torch.float32)
        if init_method is None:
            nn.init.uniform_(full_w, -sqrt(1.0 / in_size), sqrt(1.0 / in_size))
        else:
            init_method(full_w)
        full_b = torch.empty(out_size, dtype=torch.float32).uniform_(-sqrt(1.0 / out_size), sqrt(1.0 / out_size)) if bias else None
        if self.column:
            w = full_w.chunk(self.world, dim=0)[self.rank]
            b = full_b.chunk(self.world, dim=0)[self.rank] if bias else None
        else:
            w = full_w.chunk(self.world, dim=1)[self.rank]
            b = full_b
        self.weight = nn.Parameter(w.contiguous().to(dtype))
        self.bias = nn.Parameter(b.contiguous().to(dtype)) if bias else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return _linear(x, self.weight, self.bias)
        if self.column:
            y = _linear(_CopyToTP.apply(x, self.group), self.weight, self.bias)
            return _GatherLastDim.apply(y, self.group, self.rank) if self.gather_output else y
        if not self.input_is_parallel:
            x = _SplitLastDim.apply(x, self.group, self.rank)
        y = _ReduceFromTP.apply(_linear(x, self.weight, None), self.group)
        return y if self.bias is None else y + self.bias

    def extra_repr(self) -> str:
        kind = "column" if self.column else "row"
        return f"in={self.in_size}, out={self.out_size}, {kind}-parallel over {self.world}, local={tuple(self.weight.shape)}"


def make_parallel_head(n_embd: int, out: int, sequence_parallel: bool = False, dtype: torch.dtype = torch.bfloat16,
                       tp: Optional[parallel_state.ModelParallelState] = None) -> nn.Sequential:
    """``ParallelLinear(n, 2n) → ReLU → ParallelLinear(2n, out)`` (reference ``:130-149``).  When the second layer is
    row-parallel (``out < 2n``, e.g. the value head) the intermediate stays sharded between the two GEMMs — one all-reduce
    for the whole head; a vocabulary-sized head is column → column with a gathered intermediate."""
    parallel_intermediate = out < n_embd * 2
    return nn.Sequential(
        ParallelLinear(n_embd, n_embd * 2, sequence_parallel=sequence_parallel, gather_output=not parallel_intermediate,
                       dtype=dtype, tp=tp),
        nn.ReLU(),
        ParallelLinear(n_embd * 2, out, sequence_parallel=sequence_parallel, input_is_parallel=parallel_intermediate,
                       dtype=dtype, tp=tp),
    )


def shard_head_state_dict(dense: Dict[str, torch.Tensor], head: nn.Sequential) -> Dict[str, torch.Tensor]:
    """Slice the state dict of a dense ``make_head`` MLP (keys ``0.weight, 0.bias, 2.weight, 2.bias``) for a head built by
    :func:`make_parallel_head` on this rank."""
    out = {}
    for idx in ("0", "2"):
        layer: ParallelLinear = head[int(idx)]
        w, b = dense[f"{idx}.weight"], dense.get(f"{idx}.bias")
        if layer.column:
            out[f"{idx}.weight"] = w.chunk(layer.world, dim=0)[layer.rank].clone()
            if b is not None:
                out[f"{idx}.bias"] = b.chunk(layer.world, dim=0)[layer.rank].clone()
        else:
            out[f"{idx}.weight"] = w.chunk(layer.world, dim=1)[layer.rank].clone()
            if b is not None:
                out[f"{idx}.bias"] = b.clone()
    return out


class ValueHead(nn.Module):
    """Scalar value per position from a model-parallel MLP head (reference ``:152-164``): ``[B, T, H] → [B, T]``
    (``seq_first=True``: Megatron's ``[T, B, H] → [B, T]``)."""

    def __init__(self, hidden_size: int, sequence_parallel: bool = False, dtype: torch.dtype = torch.bfloat16,
                 seq_first: bool = False, tp: Optional[parallel_state.ModelParallelState] = None):
        super().__init__()
        self.hidden_size, self.sequence_parallel, self.seq_first = hidden_size, sequence_parallel, seq_first
        self.v_head = make_parallel_head(hidden_size, 1, sequence_parallel=sequence_parallel, dtype=dtype, tp=tp)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        vs = self.v_head(x).squeeze(-1)
        return vs.transpose(0, 1) if self.seq_first else vs


class RefLMHeads(PipelineTensorAPI, nn.Module):
    """Policy LM + heads with the frozen *reference* policy kept as pinned host copies of the trainable parameters and
    swapped into the same modules on demand (reference ``:167-312``) — one set of weights in HBM at any time.

    ``forward(..., run_policy_model, run_reference_model, run_value_head)`` returns ``(logits, heads_output, ref_logits)``
    like the reference; :meth:`reference` is a context manager for arbitrary code that should see the reference weights
    (generation, scoring)."""

    def __init__(self, language_model: nn.Module, other_heads: Optional[nn.Module] = None, build_reference_model: bool = True):
        super().__init__()
        self.language_model = language_model
        self.other_heads = other_heads
        self.build_reference_model = build_reference_model
        self.reference_model_offloaded = True  # True: the policy weights are the ones on the device
        self._ref: Dict[str, torch.Tensor] = {}
        self._policy: Dict[str, torch.Tensor] = {}
        if build_reference_model:
            self.snapshot_reference()

    def _trainable(self):
        return [(n, p) for n, p in self.language_model.named_parameters() if p.requires_grad]

    @torch.no_grad()
    def snapshot_reference(self) -> None:
        """Record the current trainable weights as the reference policy (called at construction and after loading a
        pretrained state dict)."""
        self._ref = {}
        for n, p in self._trainable():
            host = p.detach().to("cpu", copy=True)
            if torch.cuda.is_available():
                host = host.pin_memory()
            self._ref[n] = host

    def load_state_dict(self, state_dict, strict: bool = True, **kw):  # noqa: D401 - mirrors the reference's override
        """Load the LM weights, then re-snapshot the reference (reference ``:206-225``)."""
        keys = self.language_model.load_state_dict(state_dict, strict=strict, **kw)
        if self.build_reference_model:
            self.snapshot_reference()
        return keys

    def pretrained_state_dict(self):
        return self.language_model.state_dict()

    @torch.no_grad()
    def offload_policy_model(self) -> None:
        """Park the policy weights on the host and bring the reference weights onto the device."""
        if not self.build_reference_model or not self.reference_model_offloaded:
            return
        for n, p in self._trainable():
            self._policy[n] = p.detach().to("cpu", non_blocking=True, copy=True)
            p.data.copy_(self._ref[n].to(p.device, non_blocking=True))
        self.reference_model_offloaded = False

    @torch.no_grad()
    def offload_reference_model(self) -> None:
        """Restore the policy weights (the reference copies stay on the host)."""
        if not self.build_reference_model or self.reference_model_offloaded:
            return
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()  # the non-blocking device→host copies above must have landed
        for n, p in self._trainable():
            p.data.copy_(self._policy.pop(n).to(p.device, non_blocking=True))
        self.reference_model_offloaded = True

    @contextlib.contextmanager
    def reference(self):
        """Run the enclosed code with the reference weights in place (always restores the policy)."""
        self.offload_policy_model()
        try:
            yield self.language_model
        finally:
            self.offload_reference_model()

    def _run_lm(self, *args, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        out = self.language_model(*args, output_hidden_states=True, **kwargs)
        logits = out.logits if hasattr(out, "logits") else out[0]
        hidden = out.hidden_states[-1] if getattr(out, "hidden_states", None) is not None else None
        return logits, hidden

    def forward(self, *args, run_policy_model: bool = True, run_reference_model: bool = False, run_value_head: bool = False,
                **kwargs):
        logits = heads_output = ref_logits = None
        if run_policy_model:
            self.offload_reference_model()
            logits, hidden = self._run_lm(*args, **kwargs)
            if run_value_head and self.other_heads is not None:
                heads_output = self.other_heads(hidden)
        if run_reference_model:
            with self.reference(), torch.no_grad():
                ref_logits, _ = self._run_lm(*args, **kwargs)
        return logits, heads_output, ref_logits


_LAYER_KEY = re.compile(r"^(?P<prefix>(?:.*\.)?(?:h|layers|blocks)\.)(?P<idx>\d+)\.(?P<rest>.+)$")
_FINAL_NORM_KEY = re.compile(r"(?:^|\.)(?:ln_f|final_layernorm|final_layer_norm|final_norm|norm)\.(?:weight|bias)$")


def reshard_for_pipeline_parallelism(num_layers: int, state_dict: Dict[str, Any], pp_rank: Optional[int] = None,
                                     pp_size: Optional[int] = None) -> Dict[str, Any]:
    """Keep the transformer layers of this pipeline stage and renumber them from zero (reference ``:321-352``); the final
    norm survives only on the last stage, every other entry (embeddings, heads) is kept for the stage to pick from.
    Layer keys are recognised as ``…h.<i>.…`` / ``…layers.<i>.…`` / ``…blocks.<i>.…``."""
    st = parallel_state.get_model_parallel()
    pp_rank = st.pp_rank if pp_rank is None else pp_rank
    pp_size = st.pp_size if pp_size is None else pp_size
    stage_layers = num_layers // pp_size
    lo = pp_rank * stage_layers
    hi = num_layers if pp_rank == pp_size - 1 else lo + stage_layers
    out = {}
    for key, value in state_dict.items():
        m = _LAYER_KEY.match(key)
        if m:
            idx = int(m.group("idx"))
            if lo <= idx < hi:
                out[f"{m.group('prefix')}{idx - lo}.{m.group('rest')}"] = value
        elif _FINAL_NORM_KEY.search(key) and pp_rank != pp_size - 1:
            continue
        else:
            out[key] = value
    return out


class PPOGPT(MegatronModelMixin, nn.Module):
    """The model the model-parallel PPO trainer optimises (reference ``PPOGPT(MegatronGPTModel)``, ``:355-1222``): a
    causal LM whose blocks are sharded over the tensor-parallel group, a model-parallel :class:`ValueHead`, and the
    reference policy either as the hydra branch of the wrapped model or as :class:`RefLMHeads` host copies.

    ``PPOGPT(config)`` builds it from a :class:`~trlx_b200.data.configs.TRLConfig` with the layout published in
    :mod:`trlx_b200.parallel.state`; the NeMo-style trainer (:class:`~trlx_b200.trainer.nemo_ppo_trainer.NeMoPPOTrainer`)
    performs the same construction through ``MegatronMixin.setup_model``."""

    def __init__(self, config, build_reference_model: Optional[bool] = None, parallel_value_head: bool = True,
                 pad_token_id: int = 0, metric_fn: Optional[Callable] = None):
        super().__init__()
        self.config, self.ppo_config, self.pad_token_id, self.metric_fn = config, config.method, pad_token_id, metric_fn
        from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead, AutoModelForCausalLMWithValueHead
        from trlx_b200.parallel.tensor_parallel import apply_tensor_parallel

        st = parallel_state.get_model_parallel()
        unfrozen = config.model.num_layers_unfrozen
        hydra = unfrozen > 0
        cls = AutoModelForCausalLMWithHydraValueHead if hydra else AutoModelForCausalLMWithValueHead
        kwargs = dict(num_layers_unfrozen=unfrozen) if hydra else {}
        path = config.model.model_path
        config_like = isinstance(path, dict) or (hasattr(path, "model_type") and not isinstance(path, str)) or hasattr(path, "family")
        from_fn = cls.from_config if config_like else cls.from_pretrained
        model = from_fn(path, peft_config=config.model.peft_config, **kwargs, **(config.model.model_extra_configs or {}))
        if st.tp_group is not None and st.tp_size > 1:
            apply_tensor_parallel(model, st.tp_group, st.tp_rank, st.tp_size,
                                  sequence_parallel=bool(getattr(config.train.parallel, "sequence_parallel", False)))
        self.model = model
        hidden = model.base_model.config.hidden_size if hasattr(model.base_model, "config") else model.v_head[0].in_features
        dtype = next(model.parameters()).dtype
        if parallel_value_head and st.tp_group is not None and st.tp_size > 1:
            dense = model.v_head.state_dict()
            head = ValueHead(hidden, dtype=dtype)
            head.v_head.load_state_dict(shard_head_state_dict(dense, head.v_head))
            self.value_head: Optional[ValueHead] = head
        else:
            self.value_head = None
        if build_reference_model is None:
            build_reference_model = not hydra
        self.ref_heads = RefLMHeads(model.base_model, None, build_reference_model) if build_reference_model else None

    def forward(self, input_ids, attention_mask=None, position_ids=None, **kw):
        out = self.model(input_ids, attention_mask=attention_mask, position_ids=position_ids, output_hidden_states=True,
                         return_dict=True, **kw)
        if self.value_head is not None:
            out.value = self.value_head(out.hidden_states[-1])
        return out

    @torch.no_grad()
    def reference_logits(self, input_ids, attention_mask=None, position_ids=None):
        """Logits of the frozen reference policy for the same tokens."""
        if self.ref_heads is not None:
            with self.ref_heads.reference() as lm:
                return lm(input_ids, attention_mask=attention_mask, position_ids=position_ids).logits
        return self.model.forward_hydra(input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                        return_dict=True).logits

    def generate(self, *args, **kwargs):
        return self.model.generate(*args, **kwargs)

    # Per-rank shards are stored with the canonical (sharded) parameter names: the wrapped model's own ``state_dict`` emits
    # HF key names for whole, unsharded checkpoints, which is not what an ``mp_rank_XX`` file holds.
    def state_dict(self, *args, **kwargs):
        sd = {f"model.{k}": v for k, v in nn.Module.state_dict(self.model).items()}
        if self.value_head is not None:
            sd.update({f"value_head.{k}": v for k, v in self.value_head.state_dict().items()})
        return sd

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        res = nn.Module.load_state_dict(self.model, {k[len("model."):]: v for k, v in state_dict.items() if k.startswith("model.")},
                                        strict=strict)
        if self.value_head is not None:
            vh = {k[len("value_head."):]: v for k, v in state_dict.items() if k.startswith("value_head.")}
            if vh or strict:
                self.value_head.load_state_dict(vh, strict=strict)
        if self.ref_heads is not None and self.ref_heads.build_reference_model:
            self.ref_heads.snapshot_reference()
        return res

    def offload_reference_model(self) -> None:
        """Make sure the *policy* weights are the ones on the device (no-op with a hydra reference branch)."""
        if self.ref_heads is not None:
            self.ref_heads.offload_reference_model()

    @torch.no_grad()
    def infer_logprobs_and_values(self, input_ids, attention_mask=None, position_ids=None):
        """``(logprobs, ref_logprobs, values)`` of every next token, ``[B, T-1]`` each (reference ``:1095-1156``)."""
        from trlx_b200.utils.modeling import logprobs_of_labels

        if attention_mask is None:
            attention_mask = input_ids.ne(self.pad_token_id).long()
        if position_ids is None:
            position_ids = (attention_mask.cumsum(-1) - 1).clamp_min(0)
        with self.inference_mode():
            out = self(input_ids, attention_mask, position_ids)
            ref_logits = self.reference_logits(input_ids, attention_mask, position_ids)
        labels = input_ids[:, 1:]
        return (logprobs_of_labels(out.logits[:, :-1], labels).float(), logprobs_of_labels(ref_logits[:, :-1], labels).float(),
                out.value[:, :-1].float())

    def _loss(self, batch):
        """PPO loss of one :class:`~trlx_b200.data.ppo_types.PPORLBatch` (the fwd/loss closure of the reference, ``:945-1026``)."""
        from trlx_b200.utils.modeling import logprobs_of_labels

        method = self.ppo_config
        dev = next(self.parameters()).device
        query, response = batch.query_tensors.to(dev), batch.response_tensors.to(dev)
        old_logprobs, old_values, old_rewards = batch.logprobs.to(dev), batch.values.to(dev), batch.rewards.to(dev)
        n_resp = old_rewards.shape[1]
        advantages, returns = method.get_advantages_and_returns(old_values, old_rewards, n_resp)
        tokens = torch.cat((query, response), dim=1)
        attention_mask = tokens.ne(self.pad_token_id).long()
        position_ids = (attention_mask.cumsum(-1) - 1).clamp_min(0)
        out = self(tokens, attention_mask, position_ids)
        start, end = query.shape[1] - 1, query.shape[1] - 1 + n_resp
        logprobs = logprobs_of_labels(out.logits[:, :-1], tokens[:, 1:])[:, start:end]
        values = out.value[:, :-1][:, start:end]
        return method.loss(logprobs=logprobs.float(), values=values.float(), old_logprobs=old_logprobs, old_values=old_values,
                           advantages=advantages, returns=returns, mask=attention_mask[:, start + 1:end + 1].float())
