import os

import pytest
import torch

import trlx_b200 as trlx
from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ilql_config, default_ppo_config, default_sft_config
from trlx_b200.utils.loading import get_pipeline, get_trainer

GPT2 = dict(model_type="gpt2", vocab_size=257, n_embd=32, n_layer=3, n_head=2, n_positions=64, eos_token_id=256, bos_token_id=256)
T5 = dict(model_type="t5", vocab_size=257, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, eos_token_id=256, pad_token_id=0,
          decoder_start_token_id=0)
PROMPTS = ["hello a", "what is", "b", "count the a", "zz", "dog dog", "a dog", "the"]


def reward_fn(samples, **kwargs):
    return [float(s.count("a")) for s in samples]


def ppo_config(tmp, **over) -> TRLConfig:
    cfg = default_ppo_config().evolve(
        train=dict(seq_length=16, batch_size=2, total_steps=8, epochs=4, checkpoint_interval=4, eval_interval=4, tracker=None,
                   checkpoint_dir=str(tmp), seed=3),
        model=dict(model_path=GPT2, num_layers_unfrozen=1),
        tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(num_rollouts=4, chunk_size=2, ppo_epochs=2, gen_kwargs=dict(max_new_tokens=6, top_k=0, top_p=1.0, do_sample=True)),
    )
    return cfg.evolve(**over) if over else cfg


def test_ppo_end_to_end_checkpoints(tmp_path):
    cfg = ppo_config(tmp_path)
    trainer = trlx.train(reward_fn=reward_fn, prompts=PROMPTS, eval_prompts=PROMPTS[:2], config=cfg)
    assert trainer.iter_count == 8
    assert {"checkpoint_4", "checkpoint_8", "best_checkpoint"} <= set(os.listdir(tmp_path))
    hf = tmp_path / "checkpoint_8" / "hf_model"
    assert {"config.json", "pytorch_model.bin"} <= set(os.listdir(hf))
    keys = torch.load(hf / "pytorch_model.bin").keys()
    assert any(k.startswith("base_model.transformer.h.") for k in keys) and "v_head.0.weight" in keys
    assert any(k.startswith("frozen_head.decoder_blocks.0.") for k in keys)
    assert os.path.exists(tmp_path / "checkpoint_8" / "trainer_state_rank0.pt")
    # only the unfrozen block, final norm and value head moved
    assert all(not p.requires_grad for p in trainer.model.base_model.transformer.h[0].parameters())


@pytest.mark.parametrize("lag", [True, False])
def test_train_statistics_are_logged_once_per_step_in_order(tmp_path, lag):
    """The loop reads a step's statistics after launching the next one (``PendingStats``); every optimizer step is still logged
    exactly once, under its own index, in increasing order — checkpoints / evaluations happen on the same steps either way."""
    import json

    cfg = ppo_config(tmp_path / "ckpt", train=dict(total_steps=7, checkpoint_interval=3, eval_interval=5, tracker="jsonl",
                                                     logging_dir=str(tmp_path / "logs"), trainer_kwargs=dict(lag_stats=lag)))
    trainer = trlx.train(reward_fn=reward_fn, prompts=PROMPTS, eval_prompts=PROMPTS[:2], config=cfg)
    assert trainer.iter_count == 7
    logs = [json.loads(line) for f in os.listdir(tmp_path / "logs") for line in open(tmp_path / "logs" / f)]
    train_steps = [int(r["step"]) for r in logs if "losses/total_loss" in r]
    assert train_steps == list(range(1, 8)), train_steps
    assert all(r["losses/total_loss"] == r["losses/total_loss"] for r in logs if "losses/total_loss" in r)  # finite numbers
    assert {"checkpoint_3", "checkpoint_6", "checkpoint_7"} <= set(os.listdir(tmp_path / "ckpt"))
    evals = [int(r["step"]) for r in logs if "reward/mean" in r and int(r["step"]) > 0]
    assert evals == [5, 7], evals


def test_ppo_stats_keys_and_rollout_arithmetic(tmp_path):
    cfg = ppo_config(tmp_path, train=dict(total_steps=2, checkpoint_interval=100, eval_interval=100))
    trainer = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=reward_fn, metric_fn=None, stop_sequences=[])
    pipe = get_pipeline("PromptPipeline")(PROMPTS, 10, trainer.tokenizer)
    trainer.add_prompt_pipeline(pipe)
    trainer.make_experience(4)
    assert len(trainer.store) == 4
    for el in trainer.store.history:
        n = el.logprobs.shape[0]
        assert el.values.shape[0] == n == el.rewards.shape[0] and 1 <= n <= el.response_tensor.shape[0]
    batch = next(iter(trainer.store.create_loader(2, shuffle=False)))
    loss, stats = trainer.loss(batch)
    expected = {"losses/total_loss", "losses/policy_loss", "losses/value_loss", "values/mean", "values/clipfrac", "old_values/std",
                "returns/max", "policy/approx_kl", "policy/clipfrac", "ratio", "padding_percentage", "values/values_error",
                "values/values_mape_error"}
    assert expected <= set(stats) and torch.isfinite(loss)
    # first pass over fresh rollouts: ratio == 1, nothing clipped
    assert abs(float(stats["ratio"]) - 1) < 1e-3 and float(stats["policy/clipfrac"]) == 0


def test_ppo_dense_rewards_and_metadata(tmp_path):
    seen = {}

    def dense_reward(samples, prompts, outputs, tokenizer, **meta):
        seen.update(meta)
        return [[0.1] * max(len(tokenizer(o).input_ids), 1) for o in outputs]

    cfg = ppo_config(tmp_path, train=dict(total_steps=2, checkpoint_interval=100, eval_interval=100))
    prompts = [{"prompt": p, "tag": i} for i, p in enumerate(PROMPTS)]
    trainer = trlx.train(reward_fn=dense_reward, prompts=prompts, eval_prompts=prompts[:2], config=cfg,
                         metric_fn=lambda samples, **kw: {"len": [float(len(s)) for s in samples]})
    assert "tag" in seen and trainer.iter_count == 2


def test_ppo_with_lora_and_stop_sequences(tmp_path):
    cfg = ppo_config(tmp_path, model=dict(peft_config=dict(peft_type="LORA", r=2, lora_alpha=4)),
                     train=dict(total_steps=2, checkpoint_interval=2, eval_interval=100))
    trainer = trlx.train(reward_fn=reward_fn, prompts=PROMPTS, eval_prompts=PROMPTS[:2], config=cfg, stop_sequences=["z"])
    hf = tmp_path / "checkpoint_2" / "hf_model"
    assert {"adapter_config.json", "adapter_model.bin", "pytorch_model.bin"} <= set(os.listdir(hf))
    assert trainer.model.frozen_head is None and trainer.ref_model is None  # adapter-off forward is the reference


def test_ppo_seq2seq(tmp_path):
    cfg = ppo_config(tmp_path, model=dict(model_path=T5, model_arch_type="seq2seq", num_layers_unfrozen=1),
                     train=dict(total_steps=2, checkpoint_interval=100, eval_interval=100))
    trainer = trlx.train(reward_fn=reward_fn, prompts=PROMPTS, eval_prompts=PROMPTS[:2], config=cfg)
    assert trainer.iter_count == 2


def test_ppo_resume_from_checkpoint(tmp_path):
    cfg = ppo_config(tmp_path / "a", train=dict(total_steps=4, checkpoint_interval=4, eval_interval=100))
    t1 = trlx.train(reward_fn=reward_fn, prompts=PROMPTS, eval_prompts=PROMPTS[:2], config=cfg)
    ckpt = str(tmp_path / "a" / "checkpoint_4")
    cfg2 = ppo_config(tmp_path / "b", train=dict(total_steps=6, checkpoint_interval=100, eval_interval=100, resume_from_checkpoint=ckpt))
    t2 = get_trainer(cfg2.train.trainer)(config=cfg2, reward_fn=reward_fn, metric_fn=None, stop_sequences=[])
    t2.load(ckpt)
    for (k, a), (_, b) in zip(t1.model.raw_state_dict().items(), t2.model.raw_state_dict().items()):
        torch.testing.assert_close(a, b, msg=k)
    assert t2.iter_count == 4 and abs(t2.kl_ctl.value - t1.kl_ctl.value) < 1e-12


def test_minibatch_accumulation_equals_full_batch(tmp_path):
    def run(mb):
        torch.manual_seed(0)
        cfg = ppo_config(tmp_path / f"mb{mb}", train=dict(batch_size=4, minibatch_size=mb, total_steps=1, checkpoint_interval=100,
                                                           eval_interval=100))
        tr = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=reward_fn, metric_fn=None, stop_sequences=[])
        tr.add_prompt_pipeline(get_pipeline("PromptPipeline")(PROMPTS, 10, tr.tokenizer))
        torch.manual_seed(1)
        tr.make_experience(4)
        from trlx_b200.pipeline import MiniBatchIterator

        mbs = next(iter(MiniBatchIterator(tr.store.create_loader(4, shuffle=False), tr.mb_size, tr.num_mb)))
        assert len(mbs) == 4 // (mb or 4)
        before = tr.mb_count
        tr.train_step(mbs)
        assert tr.mb_count - before == len(mbs) and tr.iter_count == 1
        return tr

    a, b = run(None), run(2)
    assert a.num_mb == 1 and b.num_mb == 2


def test_ilql_end_to_end(tmp_path):
    cfg = default_ilql_config().evolve(
        train=dict(seq_length=16, batch_size=4, total_steps=4, epochs=4, checkpoint_interval=4, eval_interval=2, tracker=None,
                   checkpoint_dir=str(tmp_path), seed=1),
        model=dict(model_path=GPT2), tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(steps_for_target_q_sync=1, gen_kwargs=dict(max_new_tokens=4, top_k=4, beta=[0.5, 2.0], temperature=1.0)),
    )
    samples = [["ab", "cd"], ["q", "aaa"], "plain text", ["x", "yz"], ["hello", " world"], ["a", "b"], ["c", "d"], ["e", "f"]]
    rewards = [1.0, 2.0, 0.0, -1.0, 0.5, 0.1, 0.2, 0.3]
    trainer = trlx.train(samples=samples, rewards=rewards, eval_prompts=["ab", "q"], config=cfg,
                         metric_fn=lambda samples, **kw: {"reward": [float(len(s)) for s in samples]})
    assert trainer.iter_count == 4 and trainer.generate_sweep_kwarg == ("beta", [0.5, 2.0])
    keys = torch.load(tmp_path / "checkpoint_4" / "hf_model" / "pytorch_model.bin").keys()
    assert "ilql_heads.q_heads.1.2.weight" in keys and "ilql_heads.target_q_heads.0.0.bias" in keys


def test_ilql_make_experience_indices():
    from trlx_b200.trainer.accelerate_ilql_trainer import make_experience
    from trlx_b200.utils.tokenizer import build_toy_tokenizer

    tok = build_toy_tokenizer("toy://bytes")
    store = make_experience([["ab", "cd"], ["x", "y"]], [1.0, 3.0], tok, verbose=False)
    el = store[0]  # tokens: a b c d <eos> ; outputs start at index 2
    assert el.input_ids.tolist() == [97, 98, 99, 100, 256]
    assert el.actions_ixs.tolist() == [1, 2, 3] and el.states_ixs.tolist() == [1, 2, 3, 4]
    assert el.dones.tolist() == [1, 1, 1, 0] and el.rewards[:-1].abs().sum() == 0
    assert abs(store[0].rewards[-1] + store[1].rewards[-1]) < 1e-6  # standardised over the dataset


def test_sft_end_to_end(tmp_path):
    cfg = default_sft_config().evolve(
        train=dict(seq_length=16, batch_size=2, total_steps=3, epochs=3, checkpoint_interval=3, eval_interval=3, tracker=None,
                   checkpoint_dir=str(tmp_path)),
        model=dict(model_path=GPT2), tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(gen_kwargs=dict(max_new_tokens=4, do_sample=False)),
    )
    trainer = trlx.train(samples=[["ab", "cd"], ["q", "rst"], ["hello", " there"], ["x", "y"]], eval_prompts=["ab"], config=cfg)
    assert trainer.iter_count == 3
    keys = torch.load(tmp_path / "checkpoint_3" / "hf_model" / "pytorch_model.bin").keys()
    assert "transformer.wte.weight" in keys  # plain HF layout for a head-less LM
    # loss on a batch equals masked next-token cross entropy
    batch = next(iter(trainer.store.create_loader(2)))
    loss, _ = trainer.loss(batch)
    out = trainer.model.base_model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"])
    labels = batch["labels"].clone()
    labels[~batch["attention_mask"].bool()] = -100
    ref = torch.nn.functional.cross_entropy(out.logits[:, :-1].reshape(-1, 257), labels[:, 1:].reshape(-1), ignore_index=-100)
    torch.testing.assert_close(loss, ref, atol=1e-5, rtol=1e-5)


def test_rft_end_to_end(tmp_path):
    from trlx_b200.trainer.accelerate_rft_trainer import RFTConfig

    cfg = default_sft_config().evolve(
        train=dict(seq_length=16, batch_size=2, total_steps=3, epochs=4, checkpoint_interval=100, eval_interval=100, tracker=None,
                   checkpoint_dir=str(tmp_path), trainer="AccelerateRFTTrainer"),
        model=dict(model_path=GPT2), tokenizer=dict(tokenizer_path="toy://bytes"),
    )
    cfg.method = RFTConfig(name="rftconfig", gen_kwargs=dict(max_new_tokens=4, do_sample=True), start_percentile=0.5,
                           end_percentile=0.9, n_improve_steps=2, n_generations_per_prompt=3)
    trainer = trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) for s in samples], prompts=PROMPTS[:4],
                         eval_prompts=PROMPTS[:2], config=cfg)
    assert trainer.iter_count == 3 and len(trainer.generations_per_prompt) >= 1


def test_trainer_and_pipeline_registries():
    assert get_trainer("acceleratePPOtrainer") is get_trainer("AcceleratePPOTrainer")
    assert get_trainer("NeMoPPOTrainer").__name__ == "NeMoPPOTrainer"
    with pytest.raises(Exception):
        get_trainer("nope")
    with pytest.raises(Exception):
        get_pipeline("nope")


def test_train_argument_validation():
    with pytest.raises(ValueError):
        trlx.train(config=ppo_config("/tmp/x"))
    with pytest.raises(ValueError):
        trlx.train(samples=["a", "b"], rewards=[1.0], config=default_ilql_config().evolve(
            train=dict(tracker=None), model=dict(model_path=GPT2), tokenizer=dict(tokenizer_path="toy://bytes")))


def test_prompt_width_buckets():
    from trlx_b200.trainer.accelerate_ppo_trainer import _bucket_prompt_width as bucket

    assert [bucket(w) for w in (1, 8, 9, 16, 17, 33, 65, 129, 1000)] == [8, 8, 16, 16, 24, 40, 80, 160, 1024]
    assert bucket(17, 32) == 32 and bucket(32, 32) == 32
    for w in range(1, 2048):
        assert w <= bucket(w) < max(w + 8, w * 1.25 + 1)


def test_ppo_full_finetune_with_host_resident_reference(tmp_path):
    """``trainer_kwargs.offload_reference``: no second model; the reference weights are swapped in from host copies."""
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_ppo_config
    from trlx_b200.trainer.accelerate_ppo_trainer import _SwappedReference

    cfg = default_ppo_config().evolve(
        train=dict(total_steps=2, batch_size=4, seq_length=24, tracker=None, checkpoint_interval=100, eval_interval=100,
                   checkpoint_dir=str(tmp_path), trainer_kwargs=dict(offload_reference=True)),
        model=dict(model_path=dict(model_type="gpt2", vocab_size=257, n_embd=16, n_layer=2, n_head=2, n_positions=32,
                                   eos_token_id=256, bos_token_id=256), num_layers_unfrozen=-1),
        tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(num_rollouts=8, chunk_size=4, ppo_epochs=1, gen_kwargs=dict(max_new_tokens=6)))
    trainer = trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) for s in samples], prompts=["ab", "cd", "ef", "gh"] * 2,
                         eval_prompts=["ab"], config=cfg)
    assert isinstance(trainer.ref_model, _SwappedReference) and trainer.iter_count >= 2
    # the policy has moved, the reference has not: its forward differs from the policy's and restores the policy afterwards
    ids = torch.randint(0, 256, (2, 8))
    policy_before = trainer.model(ids, return_dict=True).logits.detach().clone()
    ref_logits = trainer.ref_model(ids, return_dict=True).logits
    policy_after = trainer.model(ids, return_dict=True).logits.detach()
    torch.testing.assert_close(policy_before, policy_after)
    assert (ref_logits - policy_before).abs().max() > 0


def test_divergence_watchdog_stops_training_and_keeps_checkpoints_clean(tmp_path):
    import math

    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_sft_config

    cfg = default_sft_config().evolve(
        train=dict(total_steps=50, epochs=50, batch_size=2, seq_length=16, tracker=None, checkpoint_interval=1, eval_interval=1000,
                   checkpoint_dir=str(tmp_path), trainer_kwargs=dict(max_nonfinite_steps=3)),
        model=dict(model_path=dict(model_type="gpt2", vocab_size=257, n_embd=16, n_layer=1, n_head=2, n_positions=32,
                                   eos_token_id=256, bos_token_id=256)),
        tokenizer=dict(tokenizer_path="toy://bytes"), optimizer=dict(kwargs=dict(lr=float("nan"))),
        scheduler=dict(kwargs=dict(eta_min=float("nan"))), method=dict(gen_kwargs=dict(max_new_tokens=2)))
    with pytest.raises(FloatingPointError, match="diverged"):
        trlx.train(samples=["ab", "cd", "ef", "gh"], eval_prompts=["a"], config=cfg)
    # the first step's loss is still finite (the NaN learning rate only poisons the weights afterwards): exactly that step may
    # have been checkpointed, none of the poisoned ones
    saved = [d for d in os.listdir(tmp_path) if d.startswith("checkpoint_")]
    assert len(saved) <= 1
