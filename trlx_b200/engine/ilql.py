"""ILQL generation on the CUDA engine.

The reference samples ILQL completions with a Python loop around the whole model
(``trlx/models/modeling_ilql.py:360-412``): per token one eager forward, three ``[B, V]`` head evaluations, a
``log_softmax``, a ``topk`` mask, a ``softmax`` and a ``multinomial``.  Here every decode step is one CUDA graph on the same
machinery as the PPO rollout engine — paged KV cache, the policy layer stack (persistent megakernel or the tcgen05 kernel
chain), then

* the LM head and the two *target* Q heads as fp32-output tcgen05 GEMMs (``H → V`` and ``H → 2H → V``), the value head as a
  GEMM + row-dot,
* ONE sampling kernel (``csrc/decode_ops.cu: ilql_sample_kernel``) that forms ``log π_β + β·(min(Q₁, Q₂) − V)``, applies the
  optional ``logit_mask`` row of the previous token, selects the top-k scores exactly by radix selection and draws the token
  by Gumbel-max — nothing of size ``[B, V]`` goes through PyTorch ops,
* the device-side bookkeeping kernel (EOS handling, next inputs, step counter).

``AccelerateILQLTrainer`` routes ``generate`` / ``generate_eval`` here when :meth:`ILQLDecodeEngine.why_not` returns ``None``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from trlx_b200 import ops
from trlx_b200.engine.rollout import PAGE, RolloutEngine
from trlx_b200.models.modeling_base import base_lm
from trlx_b200.utils import logging

logger = logging.get_logger(__name__)


class ILQLDecodeEngine(RolloutEngine):
    @staticmethod
    def why_not(model, gen_kwargs: Optional[Dict[str, Any]] = None, config=None, stop_sequences=None) -> Optional[str]:
        if not ops.available():
            return "the sm_100a extension is not available"
        if getattr(model, "ilql_heads", None) is None:
            return "not an ILQL model"
        if getattr(model, "peft_type", None):
            return "PEFT adapters generate through the PyTorch loop"
        from trlx_b200.parallel import state as pstate

        if pstate.get_tensor_model_parallel_world_size() > 1 or pstate.get_pipeline_model_parallel_world_size() > 1:
            return "model-parallel run"
        lm = base_lm(model.base_model)
        spec = lm.config
        if not hasattr(lm, "transformer") or lm.dtype != torch.bfloat16 or lm.device.type != "cuda":
            return "needs a decoder-only bf16 model on CUDA"
        if spec.head_dim % 8 or spec.hidden_size % 8 or spec.head_dim > 256 or spec.ffn_size % 8:
            return "head_dim / hidden / ffn sizes must be multiples of 8 (head_dim <= 256)"
        if getattr(spec, "post_norm", False) or not getattr(spec, "plain_tail", True):
            return "post-LN blocks / embedding projections (OPT-350m layout) run on the PyTorch path"
        return None

    @staticmethod
    def supports(model, gen_kwargs=None, config=None, stop_sequences=None) -> bool:
        return ILQLDecodeEngine.why_not(model, gen_kwargs, config, stop_sequences) is None

    def __init__(self, model, pad_token_id: int, eos_token_id: int, seed: int = 0, use_cuda_graph: bool = True):
        self.ilql = True
        # sampling parameters are per call (beta is routinely swept during evaluation): see `generate`
        self.beta, self.ilql_top_k, self.ilql_temperature, self.logit_mask = 1.0, 20, 1.0, None
        super().__init__(model, pad_token_id, eos_token_id, dict(do_sample=True, max_new_tokens=1), cache_trunk=False, seed=seed,
                         use_cuda_graph=use_cuda_graph)

    # ---- one decode step ------------------------------------------------------------------------------------------------
    def _decode_step(self, st):
        C, spec, lm, model = ops.C, self.spec, self.lm, self.model
        tr = lm.transformer
        x = C.embed(st["next_tokens"], st["positions"], tr.wte.weight, tr.wpe.weight if tr.wpe is not None else None,
                    spec.pos_offset, None, None)
        if tr.emb_norm is not None:
            x = C.norm(x, tr.emb_norm.weight, tr.emb_norm.bias, spec.norm_eps, spec.norm == "rmsnorm")
        if self.mega:
            x = self._mega_stack(x, st)
        else:
            for i, W in enumerate(self.layers):
                x = self._layer(x, W, st["kc"][i], st["vc"][i], st)
        hf = C.norm(x, tr.ln_f.weight, tr.ln_f.bias, spec.norm_eps, spec.norm == "rmsnorm")
        V = spec.vocab_size
        heads = model.ilql_heads
        C.gemm(hf, lm.lm_head.weight, lm.lm_head.bias, None, "none", True, st["logits"][:, :V])
        for j, head in enumerate(heads.target_q_heads):
            mid = C.gemm(hf, head[0].weight, head[0].bias, None, "relu")
            C.gemm(mid, head[2].weight, head[2].bias, None, "none", True, st["q"][j][:, :V])
        vm = C.gemm(hf, heads.v_head[0].weight, heads.v_head[0].bias, None, "relu")
        v = C.rowdot(vm, heads.v_head[2].weight.view(-1), heads.v_head[2].bias)
        two = len(heads.target_q_heads) > 1
        tok = C.ilql_sample(st["logits"], st["q"][0], st["q"][1] if two else None, v, V, float(self.beta), int(self.ilql_top_k),
                            float(self.ilql_temperature), self.seed, st["step"], st["seed_dev"], self.logit_mask,
                            st["next_tokens"] if self.logit_mask is not None else None)
        # finished rows keep emitting EOS (reference: `(1 - finished) * token + finished * eos`): the pad id handed to the
        # bookkeeping kernel is the EOS id
        C.decode_step(tok, v, None, None, st["step"], st["R"], self.eos, self.eos, st["tokens_out"], st["lp_out"], None, None,
                      st["finished"], st["resp_lens"], st["seq_lens"], st["positions"], st["next_tokens"], st["n_running"])
        st["step64"].add_(1)

    def _build_state(self, B: int, Q: int, R: int):
        st = super()._build_state(B, Q, R)
        V = self.spec.vocab_size
        f32 = dict(dtype=torch.float32, device=self.device)
        vpad = (V + 7) // 8 * 8
        st["logits"] = torch.empty(B, vpad, **f32)
        st["q"] = [torch.empty(B, vpad, **f32) for _ in self.model.ilql_heads.target_q_heads]
        return st

    # ---- public API -------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, beta=1, max_new_tokens=32, max_length=1024, temperature=1, top_k=20,
                 logit_mask=None, pad_token_id=None, eos_token_id=None, **_unused) -> torch.Tensor:
        """Drop-in for ``AutoModelForCausalLMWithILQLHeads.generate``: ``[B, Q + R]`` tokens (prompt + completion, finished
        rows padded with EOS)."""
        dev = self.device
        prompt = input_ids.to(dev, non_blocking=True)
        if eos_token_id is not None:
            self.eos = int(eos_token_id)
        mask = (attention_mask if attention_mask is not None else prompt.not_equal(self.pad)).to(dev, non_blocking=True).long()
        B, Q = prompt.shape
        R = max(min(int(max_new_tokens), int(max_length) - Q), 1)
        lm_mask = None
        if logit_mask is not None:
            lm_mask = torch.as_tensor(logit_mask).to(dev).to(torch.bool).contiguous()
        key = (float(beta), int(top_k), float(temperature), None if lm_mask is None else (lm_mask.data_ptr(), tuple(lm_mask.shape)))
        self.beta, self.ilql_top_k, self.ilql_temperature, self.logit_mask = float(beta), int(top_k), float(temperature), lm_mask
        st = self._ensure_state(B, Q, R)
        if st.get("ilql_key") != key:  # sampling parameters are baked into the captured graph
            st["graph"], st["ilql_key"] = None, key
        q_lens = mask.sum(1)
        self._reset(st, q_lens, prompt[:, -1])
        self.calls += 1
        st["seed_dev"].fill_((self.calls * 0x9E3779B1) & 0x7FFFFFFFFFFF)
        self._prefill_kv(st, prompt, mask)
        if self.use_cuda_graph:
            if st["graph"] is None:
                snap = {k: st[k].clone() for k in ("seq_lens", "positions", "next_tokens", "finished", "resp_lens", "step",
                                                   "step64", "n_running")}
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                before = ops.launch_count()
                with torch.cuda.stream(s):
                    self._decode_step(st)
                self.launches_per_step = ops.launch_count() - before
                torch.cuda.current_stream().wait_stream(s)
                for k, v in snap.items():
                    st[k].copy_(v)
                st["tokens_out"].fill_(self.eos)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph), self._static_weights():
                    self._decode_step(st)
                for k, v in snap.items():
                    st[k].copy_(v)
                st["graph"] = graph
            st["tokens_out"].fill_(self.eos)
            for s_i in range(R):
                st["graph"].replay()
                ops.add_launches(self.launches_per_step)
                if (s_i & 7) == 7 and s_i + 1 < R and int(st["n_running"].item()) == 0:
                    break
        else:
            st["tokens_out"].fill_(self.eos)
            for s_i in range(R):
                self._decode_step(st)
                if (s_i & 7) == 7 and s_i + 1 < R and int(st["n_running"].item()) == 0:
                    break
        r_max = max(int(st["resp_lens"].max().item()), 1)
        return torch.cat([prompt, st["tokens_out"][:, :r_max]], 1)

    @torch.no_grad()
    def _prefill_kv(self, st, prompt: torch.Tensor, mask: torch.Tensor):
        """Forward the first Q-1 prompt tokens and scatter their K/V into the paged caches."""
        C, lm, spec = ops.C, self.lm, self.spec
        B, Q = prompt.shape
        T = Q - 1
        if T == 0:
            return
        ids, am = prompt[:, :T], mask[:, :T]
        pos = (am.long().cumsum(-1) - 1).clamp_min(0)
        out = lm(input_ids=ids, attention_mask=am, position_ids=pos, use_cache=True, compute_logits=False)
        first = (Q - mask.long().sum(1)).to(torch.int32)
        lens = (mask.long().sum(1) - 1).to(torch.int32)
        for i, (k, v) in enumerate(out.past_key_values):
            k2 = k.transpose(1, 2).reshape(B, T, -1).contiguous()
            v2 = v.transpose(1, 2).reshape(B, T, -1).contiguous()
            C.paged_kv_write(k2, v2, st["kc"][i], st["vc"][i], st["block_table"], first, lens, spec.num_kv_heads, spec.head_dim)
