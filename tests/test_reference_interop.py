"""Interoperability with the UNMODIFIED reference (``baseline/_ref``, run in a separate interpreter):

* checkpoints — a hydra value-head model saved by the reference's ``save_pretrained`` loads here and produces the same logits,
  values and frozen-branch logits; a checkpoint saved here loads in the reference and reproduces them too (the
  checkpoint format is part of the public contract);
* data — ``tokenize_dialogue`` and the offline (ILQL) experience construction give identical token / index / reward tensors.

Skipped when the reference is not installed."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "trlx")), reason="baseline/_ref is not installed")

HEADER = textwrap.dedent("""
    import os, sys, torch
    import transformers, transformers.modeling_utils, transformers.generation  # before the stand-in packages are importable
    sys.path.insert(0, {root!r})
    sys.path.insert(0, os.path.join({root!r}, "baseline"))
    from run_reference import prepare_assets
    sys.path.remove({root!r}); sys.path.remove(os.path.join({root!r}, "baseline"))
    for m in [k for k in sys.modules if k == "trlx_b200" or k.startswith("trlx_b200.")]:
        sys.modules.pop(m)  # (prepare_assets used this framework's tokenizer builder; the reference must not see it)
    sys.path.insert(0, {shims!r}); sys.path.insert(0, {ref!r})
    import transformers_compat  # noqa: F401
    import trlx
    assert {ref!r} in trlx.__file__, trlx.__file__
""")

STAGE1 = HEADER + textwrap.dedent("""
    from trlx.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    from trlx.pipeline.offline_pipeline import tokenize_dialogue
    from trlx.trainer.accelerate_ilql_trainer import make_experience
    work = {work!r}
    model_dir, tok_dir = prepare_assets(work, tiny=True)
    torch.manual_seed(5)
    model = AutoModelForCausalLMWithHydraValueHead.from_pretrained(model_dir, num_layers_unfrozen=2).eval()
    with torch.no_grad():
        for p in model.v_head.parameters():
            p.copy_(torch.randn_like(p) * 0.1)
        for p in model.frozen_head.parameters():   # make the frozen branch distinguishable from the live one
            p.add_(torch.randn_like(p) * 0.01)
    ids = torch.load({ids!r})
    mask = torch.ones_like(ids); mask[0, :3] = 0
    pos = (mask.cumsum(-1) - 1).clamp_min(0)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True)
        hydra = model.forward_hydra(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True).logits
    model.save_pretrained(os.path.join(work, "ref_ckpt"))
    tok = transformers.AutoTokenizer.from_pretrained(tok_dir)
    dialogues = {dialogues!r}
    toks = [[(m.is_output, list(m.tokens)) for m in tokenize_dialogue(d, tok, L)] for d, L in dialogues]
    store = make_experience([d for d, _ in dialogues], {rewards!r}, tok, max_length=24, verbose=False)
    cols = {{k: [t.clone() for t in getattr(store, k)] for k in ("input_ids", "attention_mask", "rewards", "states_ixs", "actions_ixs", "dones")}}
    # value BRANCH (a trainable copy of the top block ending in the value MLP) next to the hydra policy branch
    torch.manual_seed(8)
    vb = AutoModelForCausalLMWithHydraValueHead.from_pretrained(model_dir, num_layers_unfrozen=2, num_value_layers_unfrozen=1).eval()
    with torch.no_grad():
        for p in vb.v_head.parameters():
            p.add_(torch.randn_like(p) * 0.05)
        vb_out = vb(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True)
    vb.save_pretrained(os.path.join(work, "ref_vb_ckpt"))
    # ILQL heads model: forward outputs + checkpoint
    from trlx.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads
    torch.manual_seed(6)
    ilql = AutoModelForCausalLMWithILQLHeads.from_pretrained(model_dir, two_qs=True, alpha=0.5).eval()
    with torch.no_grad():
        for p in ilql.ilql_heads.parameters():
            p.copy_(torch.randn_like(p) * 0.05)
    s_ix = torch.tensor([[4, 5, 6, 7]] * ids.shape[0]); a_ix = s_ix[:, :-1]
    with torch.no_grad():
        il_logits, qs, tqs, vs, _ = ilql(input_ids=ids, attention_mask=mask, position_ids=pos, states_ixs=s_ix, actions_ixs=a_ix)
    ilql.save_pretrained(os.path.join(work, "ref_ilql_ckpt"))
    # advantage-shifted decoding, made deterministic by top_k = 1 (argmax of log pi + beta (min Q - V)); beta = 0 -> plain greedy
    gen = {{}}
    for beta in (0.0, 1.0, 4.0):
        with torch.no_grad():
            gen[beta] = ilql.generate(input_ids=ids[:, :6], attention_mask=mask[:, :6], beta=beta, top_k=1, temperature=1.0,
                                      max_new_tokens=6, pad_token_id=1023, eos_token_id=1023)
    # PPO store collation
    from trlx.data.ppo_types import PPORLElement
    from trlx.pipeline.ppo_pipeline import ppo_collate_fn
    g = torch.Generator().manual_seed(9)
    elems = [PPORLElement(torch.randint(1, 50, (q,), generator=g), torch.randint(1, 50, (r,), generator=g),
                          torch.randn(r, generator=g), torch.randn(r, generator=g), torch.randn(r, generator=g))
             for q, r in ((3, 5), (6, 2), (4, 4))]
    collated = {{side: [getattr(ppo_collate_fn(side, 0, elems), f) for f in
                        ("query_tensors", "response_tensors", "logprobs", "values", "rewards")] for side in ("left", "right")}}
    from trlx.pipeline.offline_pipeline import PromptPipeline, DialogStore
    pp_items = {pp_prompts!r}
    tok.pad_token, tok.padding_side, tok.truncation_side = tok.eos_token, "left", "right"   # what the trainers set up
    pipe = PromptPipeline(pp_items, 5, tok)
    pp_rows = [dict(pipe[i]) for i in range(len(pipe))]
    pp_batch = dict(next(iter(pipe.create_loader(4, shuffle=False))))
    ds = DialogStore([tokenize_dialogue(d, tok, L) for d, L in dialogues], tok)
    ds_batch = dict(next(iter(ds.create_loader(4, shuffle=False))))
    # micro-batching of an optimizer batch (PPO dataclass batches and BatchEncoding dicts, uneven tail)
    from trlx.pipeline import MiniBatchIterator
    from trlx.pipeline.ppo_pipeline import PPORolloutStorage
    store = PPORolloutStorage(0, "left")
    store.clear_history()   # (the reference's store starts with a `None` placeholder that its trainer clears first)
    store.push(elems)
    ppo_mbs = [[[getattr(mb, f) for f in ("query_tensors", "response_tensors", "logprobs", "values", "rewards")] for mb in group]
               for group in MiniBatchIterator(store.create_loader(3, shuffle=False), 2, 2)]
    enc_mbs = [[dict(mb) for mb in group] for group in MiniBatchIterator(ds.create_loader(4, shuffle=False), 3, 2)]
    torch.save(dict(rows=pp_rows, batch=pp_batch, dialog=ds_batch, ppo_mbs=ppo_mbs, enc_mbs=enc_mbs),
               os.path.join(work, "pipelines_ref.pt"))
    torch.save(dict(logits=out.logits, value=out.value, hydra=hydra, toks=toks, store=cols, model_dir=model_dir, tok_dir=tok_dir,
                    vb=dict(logits=vb_out.logits, value=vb_out.value),
                    ilql=dict(logits=il_logits, qs=qs, tqs=tqs, vs=vs, gen=gen), elems=[tuple(e.__dict__.values()) for e in elems],
                    collated=collated), os.path.join(work, "stage1.pt"))
""")

ROLLOUT = textwrap.dedent("""
    from trlx.data.default_configs import default_ppo_config
    from trlx.pipeline.offline_pipeline import PromptPipeline
    from trlx.trainer.accelerate_ppo_trainer import AcceleratePPOTrainer
    work = {work!r}
    st = torch.load(os.path.join(work, "stage1.pt"), weights_only=False)
    cfg = default_ppo_config()
    # perturbed value head and frozen branch (non-zero KL penalty); the copy re-saved by this framework, because the reference's
    # loader only looks for `pytorch_model.bin` while its own `save_pretrained` writes safetensors under transformers 5
    cfg.model.model_path = os.path.join(work, "our_ckpt")
    cfg.model.num_layers_unfrozen = 2
    cfg.tokenizer.tokenizer_path = st["tok_dir"]
    cfg.train.tracker = None
    cfg.train.seq_length, cfg.train.batch_size = 40, 4
    cfg.train.checkpoint_dir = os.path.join(work, "ckpt_ref")
    cfg.method.num_rollouts, cfg.method.chunk_size = 8, 4
    cfg.method.init_kl_coef = 0.3
    cfg.method.gen_kwargs = dict(max_new_tokens=8, do_sample=False, top_k=0, top_p=1.0)
    torch.manual_seed(0)
    trainer = AcceleratePPOTrainer(config=cfg, reward_fn=lambda samples, **kw: [float(len(s)) / 10 for s in samples],
                                   metric_fn=None, stop_sequences=[])
    trainer.add_prompt_pipeline(PromptPipeline({prompts!r}, 32, trainer.tokenizer))
    trainer.make_experience(8)
    torch.save([tuple(t.detach().float().cpu() if t.is_floating_point() else t.cpu() for t in
                      (e.query_tensor, e.response_tensor, e.logprobs, e.values, e.rewards)) for e in trainer.store.history],
               os.path.join(work, "rollouts_ref.pt"))
    tk = trainer.tokenizer
    torch.save(dict(padding_side=tk.padding_side, truncation_side=tk.truncation_side, pad_token=tk.pad_token, pad_id=tk.pad_token_id,
                    eos_id=tk.eos_token_id, bos_id=tk.bos_token_id, sep=getattr(tk, "sep_token", None),
                    gen=dict(trainer.generate_kwargs), gen_exp=dict(trainer.generate_experience_kwargs or {{}}),
                    n_trainable=sum(p.numel() for p in trainer.model.parameters() if p.requires_grad)),
               os.path.join(work, "trainer_setup_ref.pt"))
    # dense (per-token) rewards that depend on prompt metadata forwarded to the reward function
    def dense_reward(samples, prompts, outputs, tokenizer, bonus, **kw):
        return [[0.05 * b * (i + 1) for i in range(len(tokenizer(o).input_ids))] for o, b in zip(outputs, bonus)]
    trainer.reward_fn = dense_reward
    trainer.config.method.chunk_size = 8
    trainer.add_prompt_pipeline(PromptPipeline([dict(prompt=p, bonus=float(i + 1)) for i, p in enumerate({prompts!r})], 32,
                                               trainer.tokenizer))
    trainer.store.clear_history()
    trainer.make_experience(8)
    torch.save([(e.query_tensor.cpu(), e.rewards.detach().float().cpu()) for e in trainer.store.history],
               os.path.join(work, "rollouts_dense_ref.pt"))
    trainer.reward_fn = lambda samples, **kw: [float(len(s)) / 10 for s in samples]
    trainer.config.method.chunk_size = 4
    # decode(): prompt / output split, stop-sequence trimming, EOS restoration
    def build(tok):
        P = [tok(t).input_ids for t in ("the movie", "acting", "film")]
        O = [tok(t).input_ids for t in (" was really boring and long", " felt good zz very", " great")]
        O[2] = O[2] + [tok.eos_token_id]
        wp, wo = max(map(len, P)), max(map(len, O))
        pad = tok.pad_token_id
        prompts = torch.tensor([[pad] * (wp - len(p)) + p for p in P])
        outs = torch.tensor([o + [pad] * (wo - len(o)) for o in O])
        return prompts, torch.cat([prompts, outs], 1)
    dec = {{}}
    for stops in ([], ["ing", "zz"]):
        trainer.stop_sequences = stops
        pt, sm = build(trainer.tokenizer)
        dec[tuple(stops)] = [trainer.decode(pt, sm, append_eos_token=flag) for flag in (True, False)]
    trainer.stop_sequences = []
    torch.save(dec, os.path.join(work, "decode_ref.pt"))
    # reward scaling variants, one chunk of 8 (running moments / reference moments do not depend on the shuffled order then)
    scaled = {{}}
    for mode, clip in (("running", 10.0), ("ref", 0.8)):
        trainer.config.method.scale_reward, trainer.config.method.cliprange_reward = mode, clip
        trainer.config.method.chunk_size = 8
        trainer.ref_mean = trainer.ref_std = None   # the reference moments are taken from the first chunk seen: take them here
        trainer.add_prompt_pipeline(PromptPipeline({prompts!r}, 32, trainer.tokenizer))
        trainer.store.clear_history()
        trainer.make_experience(8)
        scaled[mode] = [(e.query_tensor.cpu(), e.rewards.detach().float().cpu()) for e in trainer.store.history]
        scaled[mode + "/mean_kl"] = float(trainer.mean_kl)
        scaled[mode + "/running"] = (float(trainer.running_moments.mean), float(trainer.running_moments.std))
    torch.save(scaled, os.path.join(work, "rollouts_scaled_ref.pt"))
    trainer.config.method.scale_reward, trainer.config.method.cliprange_reward, trainer.config.method.chunk_size = "ignored", 10, 4
    # evaluation: greedy generations on fixed prompts, reward + metric means; then the same with a `gen_kwargs` list (sweep)
    trainer.metric_fn = lambda samples, prompts, outputs, **kw: dict(out_len=[float(len(o)) for o in outputs],
                                                                     n_e=[float(s.count("e")) for s in samples])
    trainer.add_eval_pipeline(PromptPipeline({prompts!r}[:4], 32, trainer.tokenizer))
    trainer.eval_dataloader = trainer.eval_pipeline.create_loader(2)
    ev = trainer.evaluate()
    trainer.generate_sweep_kwarg = ("max_new_tokens", [3, 6])
    ev2 = trainer.evaluate()
    trainer.generate_sweep_kwarg = None
    keep = lambda d: {{k: float(v) for k, v in d.items() if k.startswith(("reward/", "metrics/"))}}
    torch.save(dict(plain=keep(ev), sweep=keep(ev2)), os.path.join(work, "eval_ref.pt"))
    batch = next(iter(trainer.store.create_loader(4, shuffle=False)))
    trainer.model.eval()  # (HF's GPT-2 config carries dropout 0.1, which the reference leaves on while training: not comparable)
    loss, stats = trainer.loss(batch)
    loss.backward()
    inner = trainer.accelerator.unwrap_model(trainer.model)
    torch.save(dict(batch=[getattr(batch, f) for f in ("query_tensors", "response_tensors", "logprobs", "values", "rewards")],
                    loss=loss.detach(), stats={{k: float(v) for k, v in stats.items()}},
                    g_vhead=inner.v_head[2].weight.grad.clone(), g_lnf=inner.base_model.transformer.ln_f.weight.grad.clone()),
               os.path.join(work, "ppo_loss_ref.pt"))
    # one optimizer + scheduler step with the recipe's AdamW / cosine schedule
    before = inner.v_head[2].weight.detach().clone()
    trainer.opt.step(); trainer.scheduler.step()
    extra = torch.load(os.path.join(work, "ppo_loss_ref.pt"), weights_only=False)
    extra.update(w_vhead=inner.v_head[2].weight.detach().clone(), w_lnf=inner.base_model.transformer.ln_f.weight.detach().clone(),
                 moved=float((inner.v_head[2].weight.detach() - before).abs().max()), lr=float(trainer.scheduler.get_last_lr()[0]))
    torch.save(extra, os.path.join(work, "ppo_loss_ref.pt"))
""")

OFFLINE = textwrap.dedent("""
    from trlx.data.default_configs import default_ilql_config, default_sft_config
    from trlx.trainer.accelerate_ilql_trainer import AccelerateILQLTrainer
    from trlx.trainer.accelerate_sft_trainer import AccelerateSFTTrainer
    work = {work!r}
    st = torch.load(os.path.join(work, "stage1.pt"), weights_only=False)
    samples = {samples!r}
    # ---- ILQL
    cfg = default_ilql_config()
    cfg.model.model_path = os.path.join(work, "our_ilql_ckpt")
    cfg.tokenizer.tokenizer_path = st["tok_dir"]
    cfg.train.tracker, cfg.train.seq_length, cfg.train.batch_size = None, 32, 4
    cfg.train.checkpoint_dir = os.path.join(work, "ckpt_ref_ilql")
    cfg.method.alpha = 0.5
    torch.manual_seed(0)
    tr = AccelerateILQLTrainer(config=cfg, reward_fn=None, metric_fn=None, stop_sequences=[])
    tr.make_experience(samples, {rewards!r}, 32)
    batch = next(iter(tr.store.create_loader(4)))
    tr.model.eval()
    loss, stats = tr.loss(batch)
    loss.backward()
    inner = tr.accelerator.unwrap_model(tr.model)
    out = dict(ilql=dict(loss=loss.detach(), stats={{k: float(v) for k, v in stats.items()}},
                         batch=[getattr(batch, f) for f in ("input_ids", "attention_mask", "rewards", "states_ixs", "actions_ixs", "dones")],
                         g_v=inner.ilql_heads.v_head[2].weight.grad.clone(), g_q=inner.ilql_heads.q_heads[1][0].weight.grad.clone(),
                         g_lnf=inner.base_model.transformer.ln_f.weight.grad.clone()))
    # ---- SFT
    cfg = default_sft_config()
    cfg.model.model_path = st["model_dir"]
    cfg.tokenizer.tokenizer_path = st["tok_dir"]
    cfg.train.tracker, cfg.train.seq_length, cfg.train.batch_size = None, 32, 4
    cfg.train.checkpoint_dir = os.path.join(work, "ckpt_ref_sft")
    torch.manual_seed(0)
    tr = AccelerateSFTTrainer(config=cfg, reward_fn=None, metric_fn=None, stop_sequences=[])
    tr.make_experience(samples, 32)
    batch = next(iter(tr.store.create_loader(4)))
    tr.model.eval()
    loss, stats = tr.loss(batch)
    loss.backward()
    inner = tr.accelerator.unwrap_model(tr.model)
    out["sft"] = dict(loss=loss.detach(), batch={{k: v for k, v in batch.items()}},
                      g_lnf=inner.transformer.ln_f.weight.grad.clone(), g_wte=inner.transformer.wte.weight.grad.clone())
    torch.save(out, os.path.join(work, "offline_ref.pt"))
""")

LEARN = textwrap.dedent("""
    from trlx.data.default_configs import default_ppo_config, default_ilql_config, default_sft_config
    work = {work!r}
    st = torch.load(os.path.join(work, "stage1.pt"), weights_only=False)
    import accelerate
    logged = {{}}
    def capture(kind):
        def log(self, stats, step=None, **kw):
            logged.setdefault(kind, set()).update(k for k in stats if not k.startswith("samples"))
        return log
    def tree(root):
        out = []
        for d, _, fs in os.walk(root):
            out += [os.path.relpath(os.path.join(d, f), root) for f in fs]
        return sorted(out)
    trees = {{}}
    prompts = {prompts!r}
    common = dict(tracker=None, total_steps=2, epochs=2, checkpoint_interval=1, eval_interval=1, save_best=True)
    # PPO
    cfg = default_ppo_config()
    cfg.model.model_path, cfg.model.num_layers_unfrozen = os.path.join(work, "our_ckpt"), 2
    cfg.tokenizer.tokenizer_path = st["tok_dir"]
    for k, v in dict(common, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "learn_ref_ppo")).items():
        setattr(cfg.train, k, v)
    cfg.method.num_rollouts, cfg.method.chunk_size, cfg.method.ppo_epochs = 4, 4, 1
    cfg.method.gen_kwargs = dict(max_new_tokens=6, do_sample=False, top_k=0, top_p=1.0)
    accelerate.Accelerator.log = capture("ppo")
    trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) / 10 for s in samples], prompts=prompts, eval_prompts=prompts[:2], config=cfg)
    trees["ppo"] = tree(cfg.train.checkpoint_dir)
    # ILQL
    cfg = default_ilql_config()
    cfg.model.model_path, cfg.tokenizer.tokenizer_path = os.path.join(work, "our_ilql_ckpt"), st["tok_dir"]
    for k, v in dict(common, seq_length=32, batch_size=4, checkpoint_dir=os.path.join(work, "learn_ref_ilql")).items():
        setattr(cfg.train, k, v)
    cfg.method.gen_kwargs = dict(max_new_tokens=4, top_k=1, beta=1, temperature=1.0)
    accelerate.Accelerator.log = capture("ilql")
    trlx.train(samples={samples!r}, rewards=[1.0, -1.0, 0.5, 2.0], eval_prompts=prompts[:2], config=cfg)
    trees["ilql"] = tree(cfg.train.checkpoint_dir)
    # SFT
    cfg = default_sft_config()
    cfg.model.model_path, cfg.tokenizer.tokenizer_path = st["model_dir"], st["tok_dir"]
    for k, v in dict(common, seq_length=32, batch_size=4, checkpoint_dir=os.path.join(work, "learn_ref_sft")).items():
        setattr(cfg.train, k, v)
    cfg.method.gen_kwargs = dict(max_new_tokens=4, do_sample=False)
    accelerate.Accelerator.log = capture("sft")
    trlx.train(samples={samples!r}, eval_prompts=prompts[:2], config=cfg)
    trees["sft"] = tree(cfg.train.checkpoint_dir)
    torch.save(dict(logged={{k: sorted(v) for k, v in logged.items()}}, trees=trees), os.path.join(work, "learn_ref.pt"))
""")

RFTSTAGE = textwrap.dedent("""
    from trlx.data.default_configs import default_sft_config
    from trlx.trainer.accelerate_rft_trainer import AccelerateRFTTrainer, RFTConfig
    from trlx.pipeline.offline_pipeline import PromptPipeline
    work = {work!r}
    st = torch.load(os.path.join(work, "stage1.pt"), weights_only=False)
    cfg = default_sft_config()
    cfg.method = RFTConfig(name="RFTConfig", gen_kwargs=dict(max_new_tokens=6, do_sample=False), start_percentile=0.5,
                           end_percentile=0.9, n_improve_steps=2, n_generations_per_prompt=2)
    cfg.model.model_path, cfg.tokenizer.tokenizer_path = st["model_dir"], st["tok_dir"]
    cfg.train.tracker, cfg.train.seq_length, cfg.train.batch_size, cfg.train.trainer = None, 32, 4, "AccelerateRFTTrainer"
    cfg.train.checkpoint_dir = os.path.join(work, "ckpt_ref_rft")
    torch.manual_seed(0)
    tr = AccelerateRFTTrainer(config=cfg, reward_fn=lambda samples, prompts, outputs, **kw: [float(len(o) % 7) for o in outputs],
                              metric_fn=None, stop_sequences=[])
    tr.add_prompt_pipeline(PromptPipeline({prompts!r}, 16, tr.tokenizer))
    tr.add_eval_pipeline(PromptPipeline({prompts!r}[:2], 16, tr.tokenizer))
    tr.epoch_count = tr.iter_count = 0
    tr.generations_per_prompt = __import__("collections").defaultdict(list)
    selected = []
    for _ in range(2):   # a growth step, then an improvement step with the raised percentile
        tr.make_experience()
        selected.append(sorted((r["input_ids"] if isinstance(r, dict) else r.input_ids) for r in [tr.store[i] for i in range(len(tr.store))]))
        tr.epoch_count += 1
    scores = {{p: [(x["output"], x["score"]) for x in v] for p, v in tr.generations_per_prompt.items()}}
    torch.save(dict(selected=selected, scores=scores), os.path.join(work, "rft_ref.pt"))
""")

T5STAGE = textwrap.dedent("""
    from trlx.models.modeling_ppo import AutoModelForSeq2SeqLMWithValueHead
    from trlx.models.modeling_ilql import AutoModelForSeq2SeqLMWithILQLHeads
    work = {work!r}
    d = os.path.join(work, "t5_hf")
    os.makedirs(d, exist_ok=True)
    # untied embeddings (flan-t5 style, what the reference's T5 examples use).  With TIED embeddings the reference's ILQL wrapper
    # applies `lm_head` to the raw decoder state (`modeling_ilql.py:570-572`), skipping T5's d_model**-0.5 rescale that HF — and this
    # framework — apply: a documented deviation (DESIGN §2), not comparable.
    cfg = transformers.T5Config(vocab_size=128, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_decoder_layers=2, num_heads=4,
                                decoder_start_token_id=0, pad_token_id=0, eos_token_id=1, tie_word_embeddings=False)
    cfg.architectures = ["T5ForConditionalGeneration"]
    torch.manual_seed(0)
    hf = transformers.T5ForConditionalGeneration(cfg)
    cfg.save_pretrained(d)
    torch.save(hf.state_dict(), os.path.join(d, "pytorch_model.bin"))
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(2, 100, (3, 7), generator=g); mask = torch.ones_like(ids); mask[1, 5:] = 0
    dec = torch.randint(2, 100, (3, 5), generator=g); dec[:, 0] = 0
    torch.manual_seed(1)
    model = AutoModelForSeq2SeqLMWithValueHead.from_pretrained(d).eval()
    with torch.no_grad():
        for p in model.v_head.parameters():
            p.copy_(torch.randn_like(p) * 0.1)
        out = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True)
    model.save_pretrained(os.path.join(work, "t5_ref_ckpt"))
    torch.manual_seed(2)
    il = AutoModelForSeq2SeqLMWithILQLHeads.from_pretrained(d, two_qs=True, alpha=0.3).eval()
    with torch.no_grad():
        for p in il.ilql_heads.parameters():
            p.copy_(torch.randn_like(p) * 0.05)
        s_ix = torch.arange(0, 5).repeat(3, 1); a_ix = s_ix[:, :-1]
        lg, qs, tqs, vs, _, _ = il(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, states_ixs=s_ix, actions_ixs=a_ix)
    il.save_pretrained(os.path.join(work, "t5_ref_ilql_ckpt"))
    torch.save(dict(ids=ids, mask=mask, dec=dec, logits=out.logits, value=out.value, il=dict(logits=lg, qs=qs, tqs=tqs, vs=vs)),
               os.path.join(work, "t5_ref.pt"))
""")

SAMPLES = [("the movie was", " really quite good"), ("i thought", " this plot felt very long and boring"), ("film", " great"),
           ("after watching the director", " acting scenes")]

PROMPTS = ["the movie was", "i thought this film", "quite", "after watching the director", "story plot acting felt very long",
           "an", "really good scenes and", "boring but"]

STAGE2 = textwrap.dedent("""
    from trlx.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead
    work = {work!r}
    model = AutoModelForCausalLMWithHydraValueHead.from_pretrained(os.path.join(work, "our_ckpt"), num_layers_unfrozen=2).eval()
    ids = torch.load({ids!r})
    mask = torch.ones_like(ids); mask[0, :3] = 0
    pos = (mask.cumsum(-1) - 1).clamp_min(0)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True)
        hydra = model.forward_hydra(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True).logits
    torch.save(dict(logits=out.logits, value=out.value, hydra=hydra), os.path.join(work, "stage2.pt"))
""")

DIALOGUES = [("hello there general", 24), (("question one", "answer one is long enough"), 24),
             (["a first prompt", "a reply", "a follow up", "the final reply of the dialogue"], 24),
             (("short", "this reply will be truncated because the budget is tiny"), 8)]
REWARDS = [1.0, -0.5, 2.0, 0.25]
PP_PROMPTS = [dict(prompt="the movie was really quite long and boring after", tag="x", score=0.0),
              dict(prompt="i thought", tag="a", score=1.5), dict(prompt="film", tag="c", score=3.0),
              dict(prompt="after watching the director scenes story plot", tag="b", score=-2.0)]


def _run(code, work):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    res = subprocess.run([sys.executable, "-c", code], cwd=work, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-4000:]


@pytest.fixture(scope="module")
def stage1(tmp_path_factory):
    work = str(tmp_path_factory.mktemp("interop"))
    ids = torch.randint(5, 900, (3, 11), generator=torch.Generator().manual_seed(2))
    torch.save(ids, os.path.join(work, "ids.pt"))
    fmt = dict(root=ROOT, shims=os.path.join(ROOT, "baseline", "shims"), ref=REF, work=work, ids=os.path.join(work, "ids.pt"),
               dialogues=DIALOGUES, rewards=REWARDS, pp_prompts=PP_PROMPTS)
    _run(STAGE1.format(**fmt), work)
    return work, fmt, ids, torch.load(os.path.join(work, "stage1.pt"), weights_only=False)


@pytest.fixture(scope="module")
def stage2(stage1):
    """Everything else the reference has to compute, in ONE more interpreter (its start-up dominates): reload of this framework's
    re-saved checkpoints, rollouts + evaluation + loss + optimizer step, offline trainers, two-step ``learn()`` runs, T5 wrappers."""
    from trlx_b200.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads

    work, fmt, ids, ref = stage1
    _our_ckpt(work)
    path = os.path.join(work, "our_ilql_ckpt")
    if not os.path.exists(os.path.join(path, "pytorch_model.bin")):
        AutoModelForCausalLMWithILQLHeads.from_pretrained(os.path.join(work, "ref_ilql_ckpt"), two_qs=True, alpha=0.5).save_pretrained(path)
    full = dict(fmt, prompts=PROMPTS, samples=SAMPLES, rewards=[1.0, -1.0, 0.5, 2.0])
    code = HEADER.format(**full) + "".join(body.format(**full) for body in (STAGE2, ROLLOUT, OFFLINE, T5STAGE, LEARN, RFTSTAGE))
    _run(code, work)
    return stage1


def _our_ckpt(work):
    """``ref_ckpt`` (written by the reference) re-saved by this framework (``pytorch_model.bin`` layout)."""
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead

    path = os.path.join(work, "our_ckpt")
    if not os.path.exists(os.path.join(path, "pytorch_model.bin")):
        model = AutoModelForCausalLMWithHydraValueHead.from_pretrained(os.path.join(work, "ref_ckpt"), num_layers_unfrozen=2)
        model.save_pretrained(path)
    return path


def _forward(model, ids):
    mask = torch.ones_like(ids)
    mask[0, :3] = 0
    pos = (mask.cumsum(-1) - 1).clamp_min(0)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True)
        hydra = model.forward_hydra(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=True).logits
    return out.logits, out.value, hydra, mask.bool()


def test_checkpoints_round_trip_between_the_reference_and_this_framework(stage2):
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead

    work, fmt, ids, ref = stage2
    ours = AutoModelForCausalLMWithHydraValueHead.from_pretrained(os.path.join(work, "ref_ckpt"), num_layers_unfrozen=2).eval()
    logits, value, hydra, m = _forward(ours, ids)
    assert (logits - ref["logits"])[m].abs().max() < 2e-4
    assert (value - ref["value"])[m].abs().max() < 2e-4
    assert (hydra - ref["hydra"])[m].abs().max() < 2e-4
    assert (hydra - logits)[m].abs().max() > 1e-3  # the frozen branch really is a different set of weights
    assert os.path.exists(os.path.join(work, "our_ckpt", "pytorch_model.bin"))  # written by this framework (fixture), reloaded there
    back = torch.load(os.path.join(work, "stage2.pt"), weights_only=False)
    for k in ("logits", "value", "hydra"):
        assert (back[k] - ref[k])[m].abs().max() < 2e-4, k


def test_dialogue_tokenisation_and_offline_experience_match_the_reference(stage1):
    import transformers

    from trlx_b200.pipeline.offline_pipeline import tokenize_dialogue
    from trlx_b200.trainer.accelerate_ilql_trainer import make_experience

    work, fmt, ids, ref = stage1
    tok = transformers.AutoTokenizer.from_pretrained(ref["tok_dir"])
    mine = [[(m.is_output, list(m.tokens)) for m in tokenize_dialogue(d, tok, L)] for d, L in DIALOGUES]
    assert mine == ref["toks"]
    store = make_experience([d for d, _ in DIALOGUES], REWARDS, tok, max_length=24, verbose=False)
    for k, want in ref["store"].items():
        got = getattr(store, k)
        assert len(got) == len(want), k
        for a, b in zip(got, want):
            torch.testing.assert_close(torch.as_tensor(a).to(b.dtype), b, atol=1e-6, rtol=1e-6, msg=lambda m: f"{k}: {m}")


def test_ilql_heads_checkpoint_and_ppo_collation_match_the_reference(stage1):
    from trlx_b200.data.ppo_types import PPORLElement
    from trlx_b200.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads
    from trlx_b200.pipeline.ppo_pipeline import ppo_collate_fn

    work, fmt, ids, ref = stage1
    model = AutoModelForCausalLMWithILQLHeads.from_pretrained(os.path.join(work, "ref_ilql_ckpt"), two_qs=True, alpha=0.5).eval()
    mask = torch.ones_like(ids)
    mask[0, :3] = 0
    pos = (mask.cumsum(-1) - 1).clamp_min(0)
    s_ix = torch.tensor([[4, 5, 6, 7]] * ids.shape[0])
    with torch.no_grad():
        logits, qs, tqs, vs, _ = model(input_ids=ids, attention_mask=mask, position_ids=pos, states_ixs=s_ix, actions_ixs=s_ix[:, :-1])
    want = ref["ilql"]
    assert (logits - want["logits"])[mask.bool()].abs().max() < 2e-4
    for got, exp in zip(list(qs) + list(tqs) + [vs], list(want["qs"]) + list(want["tqs"]) + [want["vs"]]):
        torch.testing.assert_close(got, exp, atol=2e-4, rtol=1e-4)
    for beta, exp in want["gen"].items():
        with torch.no_grad():
            got = model.generate(input_ids=ids[:, :6], attention_mask=mask[:, :6], beta=beta, top_k=1, temperature=1.0,
                                 max_new_tokens=6, pad_token_id=1023, eos_token_id=1023)
        assert got.tolist() == exp.tolist(), (beta, got.tolist(), exp.tolist())
    assert want["gen"][0.0].tolist() != want["gen"][4.0].tolist()  # the Q / V shift really changes what is decoded
    elems = [PPORLElement(*fields) for fields in ref["elems"]]
    for side in ("left", "right"):
        batch = ppo_collate_fn(side, 0, elems)
        for f, exp in zip(("query_tensors", "response_tensors", "logprobs", "values", "rewards"), ref["collated"][side]):
            torch.testing.assert_close(getattr(batch, f), exp, msg=lambda m: f"{side} {f}: {m}")


def test_ppo_experience_matches_the_reference_rollout_arithmetic(stage2):
    """Same checkpoint (distinct frozen branch and value head), same prompts, greedy decoding: every stored rollout — query,
    response, per-token log-probs, values and KL-penalised rewards with the score on the last token — equals what the
    reference's own ``make_experience`` stores (matched by prompt; both frameworks shuffle their prompt loaders)."""
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.pipeline.offline_pipeline import PromptPipeline
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    ckpt = _our_ckpt(work)
    want = torch.load(os.path.join(work, "rollouts_ref.pt"), weights_only=False)
    cfg = default_ppo_config().evolve(
        model=dict(model_path=ckpt, num_layers_unfrozen=2),
        tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours")),
        method=dict(num_rollouts=8, chunk_size=4, init_kl_coef=0.3, gen_kwargs=dict(max_new_tokens=8, do_sample=False, top_k=0, top_p=1.0)))
    torch.manual_seed(0)
    trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [float(len(s)) / 10 for s in samples],
                                             metric_fn=None, stop_sequences=[])
    trainer.add_prompt_pipeline(PromptPipeline(PROMPTS, 32, trainer.tokenizer))
    trainer.make_experience(8)
    pad = trainer.tokenizer.pad_token_id
    mine = {tuple(int(t) for t in e.query_tensor.tolist() if t != pad): e for e in trainer.store.history}
    assert len(want) == 8 == len(mine)
    for q, r, lp, v, rw in want:
        e = mine[tuple(int(t) for t in q.tolist() if t != pad)]
        # (the reference stores the response padded to its chunk's width, this framework trims it: trailing pads carry no maths —
        # log-probs / values / rewards below are sliced to the true length in both)
        def strip(t):
            t = t.tolist()
            while t and t[-1] == pad:
                t.pop()
            return t

        assert strip(e.response_tensor) == strip(r)
        assert len(e.logprobs) == len(lp) == len(v) == len(rw)
        # One documented deviation: with pad == eos the last stored position of a finished sample is a token the
        # reference's attention mask counts as padding, and the reference gives padded positions `position_id = 1`
        # (`accelerate_ppo_trainer.py:417-418`) while this framework keeps counting.  That position is masked out of the loss; its
        # log-prob / value are compared loosely, everything else — including every reward — exactly.
        n = len(lp)
        tail_is_pad = n >= 2 and int(e.response_tensor[n - 2]) == pad
        k = n - 1 if tail_is_pad else n
        torch.testing.assert_close(e.logprobs.float().cpu()[:k], lp[:k], atol=2e-4, rtol=1e-4)
        torch.testing.assert_close(e.values.float().cpu()[:k], v[:k], atol=2e-4, rtol=1e-4)
        torch.testing.assert_close(e.rewards.float().cpu(), rw, atol=2e-4, rtol=1e-4)
        assert (e.values.float().cpu()[k:] - v[k:]).abs().max().item() < 1.0 if k < n else True
        assert rw[:-1].abs().max() > 1e-4  # the KL penalty is really there (frozen branch differs from the policy)
    # dense per-token rewards driven by prompt metadata (`bonus`) that the pipeline forwards to the reward function
    def dense_reward(samples, prompts, outputs, tokenizer, bonus, **kw):
        return [[0.05 * b * (i + 1) for i in range(len(tokenizer(o).input_ids))] for o, b in zip(outputs, bonus)]

    dense = torch.load(os.path.join(work, "rollouts_dense_ref.pt"), weights_only=False)
    trainer.reward_fn = dense_reward
    trainer.config.method.chunk_size = 8
    trainer.add_prompt_pipeline(PromptPipeline([dict(prompt=p, bonus=float(i + 1)) for i, p in enumerate(PROMPTS)], 32, trainer.tokenizer))
    trainer.store.clear_history()
    trainer.make_experience(8)
    mine = {tuple(int(t) for t in e.query_tensor.tolist() if t != pad): e for e in trainer.store.history}
    assert len(dense) == 8
    for q, rw in dense:
        e = mine[tuple(int(t) for t in q.tolist() if t != pad)]
        torch.testing.assert_close(e.rewards.float().cpu(), rw, atol=3e-4, rtol=2e-4, msg=lambda m: f"dense rewards: {m}")
    trainer.reward_fn = lambda samples, **kw: [float(len(s)) / 10 for s in samples]
    # reward scaling by running / reference moments and reward clipping (single chunk: order-independent statistics)
    scaled = torch.load(os.path.join(work, "rollouts_scaled_ref.pt"), weights_only=False)
    for mode, clip in (("running", 10.0), ("ref", 0.8)):
        trainer.config.method.scale_reward, trainer.config.method.cliprange_reward = mode, clip
        trainer.config.method.chunk_size = 8
        trainer.ref_mean = trainer.ref_std = None
        trainer.add_prompt_pipeline(PromptPipeline(PROMPTS, 32, trainer.tokenizer))
        trainer.store.clear_history()
        trainer.make_experience(8)
        mine = {tuple(int(t) for t in e.query_tensor.tolist() if t != pad): e for e in trainer.store.history}
        for q, rw in scaled[mode]:
            e = mine[tuple(int(t) for t in q.tolist() if t != pad)]
            torch.testing.assert_close(e.rewards.float().cpu(), rw, atol=3e-4, rtol=2e-4, msg=lambda m: f"scale_reward={mode}: {m}")
        # the mean KL the adaptive controller is fed (k3 estimator summed over the response, averaged over the chunk)
        assert abs(float(trainer.mean_kl) - scaled[mode + "/mean_kl"]) < 2e-4 * max(1.0, abs(scaled[mode + "/mean_kl"])), mode


def _stats_close(mine, want, tol=2e-4):
    assert set(want) <= set(mine), set(want) - set(mine)
    for k, v in want.items():
        assert abs(float(mine[k]) - v) <= tol * max(1.0, abs(v)), (k, float(mine[k]), v)


def test_ppo_trainer_loss_and_gradients_match_the_reference(stage2):
    """``trainer.loss(batch)`` on the reference's own collated batch: loss, all statistics, and gradients."""
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.data.ppo_types import PPORLBatch
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    ckpt = _our_ckpt(work)
    want = torch.load(os.path.join(work, "ppo_loss_ref.pt"), weights_only=False)
    cfg = default_ppo_config().evolve(
        model=dict(model_path=ckpt, num_layers_unfrozen=2), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours2"),
                   trainer_kwargs=dict(cache_trunk=False)),
        method=dict(num_rollouts=8, chunk_size=4, init_kl_coef=0.3, gen_kwargs=dict(max_new_tokens=8, do_sample=False, top_k=0, top_p=1.0)))
    trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [0.0] * len(samples), metric_fn=None,
                                             stop_sequences=[])
    trainer.model.eval()
    loss, stats = trainer.loss(PPORLBatch(*want["batch"]))
    loss.backward()
    torch.testing.assert_close(loss.detach().float().cpu(), want["loss"], atol=2e-5, rtol=1e-4)
    _stats_close({k: float(v) for k, v in stats.items()}, want["stats"])
    torch.testing.assert_close(trainer.model.v_head[2].weight.grad, want["g_vhead"], atol=1e-5, rtol=1e-3)
    torch.testing.assert_close(trainer.model.base_model.transformer.ln_f.weight.grad, want["g_lnf"], atol=1e-5, rtol=1e-3)
    # the optimizer step of the default recipe (AdamW 3e-5, betas (0.9, 0.95), weight decay 1e-6, cosine schedule) moves the weights
    # exactly as the reference's does
    trainer.opt.step()
    trainer.scheduler.step()
    assert want["moved"] > 1e-6
    torch.testing.assert_close(trainer.model.v_head[2].weight.detach(), want["w_vhead"], atol=1e-7, rtol=1e-5)
    torch.testing.assert_close(trainer.model.base_model.transformer.ln_f.weight.detach(), want["w_lnf"], atol=1e-7, rtol=1e-5)
    assert abs(float(trainer.scheduler.get_last_lr()[0]) - want["lr"]) < 1e-12


def test_ilql_and_sft_trainer_losses_and_gradients_match_the_reference(stage2):
    from trlx_b200.data.default_configs import default_ilql_config, default_sft_config
    from trlx_b200.data.ilql_types import ILQLBatch
    from trlx_b200.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    path = os.path.join(work, "our_ilql_ckpt")
    want = torch.load(os.path.join(work, "offline_ref.pt"), weights_only=False)
    # ---- ILQL: same stored experience, loss on the reference's batch
    cfg = default_ilql_config().evolve(
        model=dict(model_path=path), tokenizer=dict(tokenizer_path=ref["tok_dir"]), method=dict(alpha=0.5),
        train=dict(tracker=None, seq_length=32, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours_ilql")))
    tr = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=None, metric_fn=None, stop_sequences=[])
    tr.make_experience(SAMPLES, [1.0, -1.0, 0.5, 2.0], 32)  # the trainer's own experience: same first batch as the reference's
    mine = next(iter(tr.store.create_loader(4)))
    fields = ("input_ids", "attention_mask", "rewards", "states_ixs", "actions_ixs", "dones")

    def rows(cols):  # the loaders shuffle: compare the batch as a set of rows
        return sorted(tuple(tuple(c[i].reshape(-1).tolist()) for c in cols) for i in range(len(cols[0])))

    got_rows, exp_rows = rows([getattr(mine, f).cpu() for f in fields]), rows(want["ilql"]["batch"])
    for g, e in zip(got_rows, exp_rows):
        for f, a, b in zip(fields, g, e):
            assert len(a) == len(b) and all(abs(x - y) < 1e-6 for x, y in zip(a, b)), f"ILQL store {f}"
    tr.model.eval()
    loss, stats = tr.loss(ILQLBatch(*want["ilql"]["batch"]))
    loss.backward()
    torch.testing.assert_close(loss.detach().float().cpu(), want["ilql"]["loss"], atol=1e-4, rtol=1e-4)
    _stats_close({k: float(v) for k, v in stats.items()}, want["ilql"]["stats"], tol=5e-4)
    heads = tr.model.ilql_heads
    torch.testing.assert_close(heads.v_head[2].weight.grad, want["ilql"]["g_v"], atol=1e-5, rtol=2e-3)
    torch.testing.assert_close(heads.q_heads[1][0].weight.grad, want["ilql"]["g_q"], atol=1e-5, rtol=2e-3)
    torch.testing.assert_close(tr.model.base_model.transformer.ln_f.weight.grad, want["ilql"]["g_lnf"], atol=1e-5, rtol=2e-3)
    # ---- SFT
    cfg = default_sft_config().evolve(
        model=dict(model_path=ref["model_dir"]), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=32, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours_sft")))
    tr = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=None, metric_fn=None, stop_sequences=[])
    tr.make_experience(SAMPLES, 32)
    mine = dict(next(iter(tr.store.create_loader(4))))
    assert set(mine) == set(want["sft"]["batch"])
    keys = sorted(mine)
    assert rows([torch.as_tensor(mine[k]).cpu() for k in keys]) == rows([want["sft"]["batch"][k] for k in keys]), "SFT store"
    tr.model.eval()
    from transformers import BatchEncoding

    enc = BatchEncoding(dict(want["sft"]["batch"]))
    loss, stats = tr.loss(enc)
    lm = tr.model.base_model if hasattr(tr.model, "base_model") else tr.model
    if bool((enc["attention_mask"][:, 0] == 0).any()):
        # Documented deviation: the reference calls HF's forward without position ids, so LEFT-padded rows are
        # embedded at positions shifted by their padding; here positions always follow the attention mask (as HF's own
        # `generate` does).  With the reference's positions this model gives the reference's loss and gradients exactly.
        assert abs(float(loss.detach()) - float(want["sft"]["loss"])) < 0.2
        ids, mask = enc["input_ids"], enc["attention_mask"]
        labels = enc["labels"].clone() if "labels" in enc else ids.clone()  # DialogStore masks the prompt tokens out
        labels[~mask.bool()] = -100
        lm.zero_grad()
        loss = lm(input_ids=ids, attention_mask=mask, position_ids=torch.arange(ids.shape[1]).expand_as(ids), labels=labels).loss
    loss.backward()
    torch.testing.assert_close(loss.detach().float().cpu(), want["sft"]["loss"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lm.transformer.ln_f.weight.grad, want["sft"]["g_lnf"], atol=1e-5, rtol=2e-3)
    torch.testing.assert_close(lm.transformer.wte.weight.grad, want["sft"]["g_wte"], atol=1e-5, rtol=2e-3)


def test_evaluate_statistics_match_the_reference(stage2):
    """``evaluate()``: greedy generations on fixed prompts → identical ``reward/mean`` and ``metrics/*`` means, with and without a
    generation-kwarg sweep (``reward/mean@max_new_tokens=…`` keys)."""
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.pipeline.offline_pipeline import PromptPipeline
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    ckpt = _our_ckpt(work)
    want = torch.load(os.path.join(work, "eval_ref.pt"), weights_only=False)
    cfg = default_ppo_config().evolve(
        model=dict(model_path=ckpt, num_layers_unfrozen=2), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours3")),
        method=dict(num_rollouts=8, chunk_size=4, gen_kwargs=dict(max_new_tokens=8, do_sample=False, top_k=0, top_p=1.0)))
    trainer = get_trainer(cfg.train.trainer)(
        config=cfg, reward_fn=lambda samples, **kw: [float(len(s)) / 10 for s in samples],
        metric_fn=lambda samples, prompts, outputs, **kw: dict(out_len=[float(len(o)) for o in outputs],
                                                                n_e=[float(s.count("e")) for s in samples]), stop_sequences=[])
    trainer.add_eval_pipeline(PromptPipeline(PROMPTS[:4], 32, trainer.tokenizer))
    trainer.eval_dataloader = trainer.eval_pipeline.create_loader(2)
    keep = lambda d: {k: float(v) for k, v in d.items() if k.startswith(("reward/", "metrics/"))}  # noqa: E731
    plain = keep(trainer.evaluate())
    trainer.generate_sweep_kwarg = ("max_new_tokens", [3, 6])
    sweep = keep(trainer.evaluate())
    assert set(plain) == set(want["plain"]) and set(sweep) == set(want["sweep"]), (set(plain), set(want["plain"]), set(sweep), set(want["sweep"]))
    for got, exp in ((plain, want["plain"]), (sweep, want["sweep"])):
        for k, v in exp.items():
            assert abs(got[k] - v) < 1e-6 * max(1.0, abs(v)), (k, got[k], v)


def test_seq2seq_value_head_and_ilql_checkpoints_from_the_reference_load_here(stage2):
    """T5: value-head and ILQL-heads wrappers saved by the reference load here with the same logits / values / Q values.  (The
    reference's seq2seq *hydra* branch is not compared: under this image's transformers its frozen branch no longer reproduces the
    model it was copied from.)"""
    from trlx_b200.models.modeling_ilql import AutoModelForSeq2SeqLMWithILQLHeads
    from trlx_b200.models.modeling_ppo import AutoModelForSeq2SeqLMWithValueHead

    work, fmt, _, _ = stage2
    ref = torch.load(os.path.join(work, "t5_ref.pt"), weights_only=False)
    ids, mask, dec = ref["ids"], ref["mask"], ref["dec"]
    model = AutoModelForSeq2SeqLMWithValueHead.from_pretrained(os.path.join(work, "t5_ref_ckpt")).eval()
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True)
    torch.testing.assert_close(out.logits, ref["logits"], atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(out.value, ref["value"], atol=2e-4, rtol=1e-4)
    il = AutoModelForSeq2SeqLMWithILQLHeads.from_pretrained(os.path.join(work, "t5_ref_ilql_ckpt"), two_qs=True, alpha=0.3).eval()
    s_ix = torch.arange(0, 5).repeat(3, 1)
    with torch.no_grad():
        lg, qs, tqs, vs, _, _ = il(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, states_ixs=s_ix, actions_ixs=s_ix[:, :-1])
    torch.testing.assert_close(lg, ref["il"]["logits"], atol=2e-4, rtol=1e-4)
    for got, exp in zip(list(qs) + list(tqs) + [vs], list(ref["il"]["qs"]) + list(ref["il"]["tqs"]) + [ref["il"]["vs"]]):
        torch.testing.assert_close(got, exp, atol=2e-4, rtol=1e-4)


def test_learn_logs_the_same_statistic_keys_and_writes_the_same_checkpoint_tree(stage2, monkeypatch):
    """Two optimizer steps of ``trlx.train`` per method in both frameworks: every statistic key the reference hands to its tracker
    is logged here too (dashboards keep working), and the checkpoint directory has the same sub-directories (``checkpoint_N``,
    ``best_checkpoint``) with an ``hf_model`` folder holding ``config.json`` + weights."""
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_ilql_config, default_ppo_config, default_sft_config
    from trlx_b200.models.modeling_ilql import AutoModelForCausalLMWithILQLHeads
    from trlx_b200.parallel.runtime import Runtime

    work, fmt, ids, ref = stage2
    path = os.path.join(work, "our_ilql_ckpt")
    want = torch.load(os.path.join(work, "learn_ref.pt"), weights_only=False)

    logged = {}
    kind = {"name": None}
    monkeypatch.setattr(Runtime, "log", lambda self, stats, step=None: logged.setdefault(kind["name"], set()).update(stats))
    common = dict(tracker=None, total_steps=2, epochs=2, checkpoint_interval=1, eval_interval=1, save_best=True)

    def tree(root):
        out = []
        for d, _, fs in os.walk(root):
            out += [os.path.relpath(os.path.join(d, f), root) for f in fs]
        return sorted(out)

    trees = {}
    kind["name"] = "ppo"
    cfg = default_ppo_config().evolve(
        model=dict(model_path=os.path.join(work, "our_ckpt"), num_layers_unfrozen=2), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(common, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "learn_our_ppo")),
        method=dict(num_rollouts=4, chunk_size=4, ppo_epochs=1, gen_kwargs=dict(max_new_tokens=6, do_sample=False, top_k=0, top_p=1.0)))
    trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) / 10 for s in samples], prompts=PROMPTS, eval_prompts=PROMPTS[:2], config=cfg)
    trees["ppo"] = tree(cfg.train.checkpoint_dir)
    kind["name"] = "ilql"
    cfg = default_ilql_config().evolve(
        model=dict(model_path=path), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(common, seq_length=32, batch_size=4, checkpoint_dir=os.path.join(work, "learn_our_ilql")),
        method=dict(gen_kwargs=dict(max_new_tokens=4, top_k=1, beta=1, temperature=1.0)))
    trlx.train(samples=SAMPLES, rewards=[1.0, -1.0, 0.5, 2.0], eval_prompts=PROMPTS[:2], config=cfg)
    trees["ilql"] = tree(cfg.train.checkpoint_dir)
    kind["name"] = "sft"
    cfg = default_sft_config().evolve(
        model=dict(model_path=ref["model_dir"]), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(common, seq_length=32, batch_size=4, checkpoint_dir=os.path.join(work, "learn_our_sft")),
        method=dict(gen_kwargs=dict(max_new_tokens=4, do_sample=False)))
    trlx.train(samples=SAMPLES, eval_prompts=PROMPTS[:2], config=cfg)
    trees["sft"] = tree(cfg.train.checkpoint_dir)

    for k in ("ppo", "ilql", "sft"):
        missing = set(want["logged"][k]) - logged[k]
        assert not missing, (k, sorted(missing))
        ref_dirs = {p.split(os.sep)[0] for p in want["trees"][k]}
        our_dirs = {p.split(os.sep)[0] for p in trees[k]}
        assert ref_dirs <= our_dirs, (k, ref_dirs, our_dirs)
        for d in ref_dirs:
            ours_hf = {os.path.basename(p) for p in trees[k] if p.startswith(os.path.join(d, "hf_model") + os.sep)}
            assert "config.json" in ours_hf and ({"pytorch_model.bin", "model.safetensors"} & ours_hf), (k, d, ours_hf)
            ref_hf = {os.path.basename(p) for p in want["trees"][k] if p.startswith(os.path.join(d, "hf_model") + os.sep)}
            # (weights may be .bin or .safetensors; `generation_config.json` is HF's own addition for its model classes)
            assert ref_hf - {"pytorch_model.bin", "model.safetensors", "generation_config.json"} <= ours_hf, (k, d, ref_hf, ours_hf)


def test_prompt_pipeline_and_dialog_store_batches_match_the_reference(stage1):
    """Prompt truncation to ``max_prompt_length``, metadata pass-through, left-padded collation; dialogue store collation with the
    loss mask of the non-output tokens."""
    import transformers

    from trlx_b200.pipeline.offline_pipeline import DialogStore, PromptPipeline, tokenize_dialogue

    work, fmt, ids, ref = stage1
    want = torch.load(os.path.join(work, "pipelines_ref.pt"), weights_only=False)
    tok = transformers.AutoTokenizer.from_pretrained(ref["tok_dir"])
    tok.pad_token, tok.padding_side, tok.truncation_side = tok.eos_token, "left", "right"
    pipe = PromptPipeline(PP_PROMPTS, 5, tok)
    assert len(pipe) == len(want["rows"])
    for i, exp in enumerate(want["rows"]):
        got = dict(pipe[i])
        assert set(got) == set(exp), (set(got), set(exp))
        for k, v in exp.items():
            assert list(got[k]) == list(v) if isinstance(v, (list, tuple)) else got[k] == v, (i, k, got[k], v)
    batch = dict(next(iter(pipe.create_loader(4, shuffle=False))))
    assert set(batch) == set(want["batch"])
    for k, v in want["batch"].items():
        if isinstance(v, torch.Tensor):
            assert torch.equal(torch.as_tensor(batch[k]), v), k
        else:
            assert list(batch[k]) == list(v), k
    ds = DialogStore([tokenize_dialogue(d, tok, L) for d, L in DIALOGUES], tok)
    got = dict(next(iter(ds.create_loader(4, shuffle=False))))
    assert set(got) == set(want["dialog"])
    for k, v in want["dialog"].items():
        assert torch.equal(torch.as_tensor(got[k]), v), k


def test_minibatch_iterator_matches_the_reference(stage1):
    from trlx_b200.data.ppo_types import PPORLElement
    from trlx_b200.pipeline import MiniBatchIterator
    from trlx_b200.pipeline.offline_pipeline import DialogStore, tokenize_dialogue
    from trlx_b200.pipeline.ppo_pipeline import PPORolloutStorage
    import transformers

    work, fmt, ids, ref = stage1
    want = torch.load(os.path.join(work, "pipelines_ref.pt"), weights_only=False)
    store = PPORolloutStorage(0, "left")
    store.clear_history()
    store.push([PPORLElement(*fields) for fields in ref["elems"]])
    got = [[[getattr(mb, f) for f in ("query_tensors", "response_tensors", "logprobs", "values", "rewards")] for mb in group]
           for group in MiniBatchIterator(store.create_loader(3, shuffle=False), 2, 2)]
    assert len(got) == len(want["ppo_mbs"])
    for g, w in zip(got, want["ppo_mbs"]):
        assert len(g) == len(w)
        for mb_g, mb_w in zip(g, w):
            for a, b in zip(mb_g, mb_w):
                torch.testing.assert_close(torch.as_tensor(a).cpu(), b)
    tok = transformers.AutoTokenizer.from_pretrained(ref["tok_dir"])
    tok.pad_token, tok.padding_side, tok.truncation_side = tok.eos_token, "left", "right"
    ds = DialogStore([tokenize_dialogue(d, tok, L) for d, L in DIALOGUES], tok)
    got = [[dict(mb) for mb in group] for group in MiniBatchIterator(ds.create_loader(4, shuffle=False), 3, 2)]
    assert len(got) == len(want["enc_mbs"])
    for g, w in zip(got, want["enc_mbs"]):
        assert len(g) == len(w)
        for mb_g, mb_w in zip(g, w):
            assert set(mb_g) == set(mb_w)
            for k in mb_w:
                assert torch.equal(torch.as_tensor(mb_g[k]), mb_w[k]), k


def test_decode_trims_stop_sequences_and_restores_eos_like_the_reference(stage2):
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    want = torch.load(os.path.join(work, "decode_ref.pt"), weights_only=False)
    cfg = default_ppo_config().evolve(
        model=dict(model_path=_our_ckpt(work), num_layers_unfrozen=2), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours4")))
    trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [0.0] * len(samples), metric_fn=None,
                                             stop_sequences=[])
    tok = trainer.tokenizer
    P = [tok(t).input_ids for t in ("the movie", "acting", "film")]
    O = [tok(t).input_ids for t in (" was really boring and long", " felt good zz very", " great")]
    O[2] = O[2] + [tok.eos_token_id]
    wp, wo, pad = max(map(len, P)), max(map(len, O)), tok.pad_token_id
    prompts = torch.tensor([[pad] * (wp - len(p)) + p for p in P])
    samples = torch.cat([prompts, torch.tensor([o + [pad] * (wo - len(o)) for o in O])], 1)
    for stops, exp in want.items():
        trainer.stop_sequences = list(stops)
        got = [trainer.decode(prompts, samples, append_eos_token=flag) for flag in (True, False)]
        assert [tuple(map(list, g)) for g in got] == [tuple(map(list, e)) for e in exp], (stops, got, exp)


def test_rft_generation_scoring_and_percentile_selection_match_the_reference(stage2):
    """RFT (greedy, so deterministic): the generations kept per prompt with their scores, and the samples that survive the
    per-prompt percentile thresholds of a growth step and of the following improvement step."""
    import collections

    from trlx_b200.data.default_configs import default_sft_config
    from trlx_b200.pipeline.offline_pipeline import PromptPipeline
    from trlx_b200.trainer.accelerate_rft_trainer import RFTConfig
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    want = torch.load(os.path.join(work, "rft_ref.pt"), weights_only=False)
    cfg = default_sft_config().evolve(
        model=dict(model_path=ref["model_dir"]), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=32, batch_size=4, trainer="AccelerateRFTTrainer", checkpoint_dir=os.path.join(work, "ckpt_ours_rft")))
    cfg.method = RFTConfig(name="RFTConfig", gen_kwargs=dict(max_new_tokens=6, do_sample=False), start_percentile=0.5,
                           end_percentile=0.9, n_improve_steps=2, n_generations_per_prompt=2)
    tr = get_trainer("AccelerateRFTTrainer")(config=cfg, reward_fn=lambda samples, prompts, outputs, **kw: [float(len(o) % 7) for o in outputs],
                                             metric_fn=None, stop_sequences=[])
    tr.add_prompt_pipeline(PromptPipeline(PROMPTS, 16, tr.tokenizer))
    tr.add_eval_pipeline(PromptPipeline(PROMPTS[:2], 16, tr.tokenizer))
    tr.epoch_count = tr.iter_count = 0
    tr.generations_per_prompt = collections.defaultdict(list)
    for step in range(2):
        tr.make_experience()
        got = sorted(list((r["input_ids"] if isinstance(r, dict) else r.input_ids)) for r in [tr.store[i] for i in range(len(tr.store))])
        assert got == [list(x) for x in want["selected"][step]], (step, got, want["selected"][step])
        tr.epoch_count += 1
    mine = {p: [(x["output"], x["score"]) for x in v] for p, v in tr.generations_per_prompt.items()}
    assert mine == want["scores"]


def test_value_branch_checkpoint_from_the_reference_loads_here(stage1):
    """``num_value_layers_unfrozen=1``: the value function is a separate copy of the top block + MLP; its checkpoint keys
    (``v_head.decoder_blocks…``) and outputs carry over."""
    from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead

    work, fmt, ids, ref = stage1
    model = AutoModelForCausalLMWithHydraValueHead.from_pretrained(os.path.join(work, "ref_vb_ckpt"), num_layers_unfrozen=2,
                                                                   num_value_layers_unfrozen=1).eval()
    logits, value, _, m = _forward(model, ids)
    assert (logits - ref["vb"]["logits"])[m].abs().max() < 2e-4
    assert (value - ref["vb"]["value"])[m].abs().max() < 2e-4
    plain = AutoModelForCausalLMWithHydraValueHead.from_pretrained(os.path.join(work, "ref_ckpt"), num_layers_unfrozen=2).eval()
    assert (_forward(plain, ids)[1] - value)[m].abs().max() > 1e-3  # a different value function than the plain head's



def test_trainer_setup_matches_the_reference(stage2):
    """What the PPO trainer derives at construction: tokenizer sides / pad token, the generation kwargs actually used for
    evaluation and for experience, and the number of trainable elements (``num_layers_unfrozen = 2`` + value head)."""
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.utils.loading import get_trainer

    work, fmt, ids, ref = stage2
    want = torch.load(os.path.join(work, "trainer_setup_ref.pt"), weights_only=False)
    cfg = default_ppo_config().evolve(
        model=dict(model_path=_our_ckpt(work), num_layers_unfrozen=2), tokenizer=dict(tokenizer_path=ref["tok_dir"]),
        train=dict(tracker=None, seq_length=40, batch_size=4, checkpoint_dir=os.path.join(work, "ckpt_ours5"),
                   trainer_kwargs=dict(cache_trunk=False)),
        method=dict(num_rollouts=8, chunk_size=4, init_kl_coef=0.3, gen_kwargs=dict(max_new_tokens=8, do_sample=False, top_k=0, top_p=1.0)))
    os.environ["TRLX_B200_SHARE_TRUNK"] = "0"  # compare the reference's trainable set (it differentiates through the whole trunk)
    try:
        trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [0.0] * len(samples), metric_fn=None,
                                                 stop_sequences=[])
    finally:
        os.environ.pop("TRLX_B200_SHARE_TRUNK", None)
    tk = trainer.tokenizer
    got = dict(padding_side=tk.padding_side, truncation_side=tk.truncation_side, pad_token=tk.pad_token, pad_id=tk.pad_token_id,
               eos_id=tk.eos_token_id, bos_id=tk.bos_token_id, sep=getattr(tk, "sep_token", None))
    for k, v in got.items():
        assert v == want[k], (k, v, want[k])
    for name, mine in (("gen", trainer.generate_kwargs), ("gen_exp", trainer.generate_experience_kwargs or {})):
        for k, v in want[name].items():
            if k == "synced_gpus":  # DeepSpeed ZeRO-3 plumbing of HF's generate; ZeRO-3 is handled inside this framework's sampler
                continue
            assert k in mine and mine[k] == v, (name, k, mine.get(k), v)
    assert sum(p.numel() for p in trainer.model.parameters() if p.requires_grad) == want["n_trainable"]
