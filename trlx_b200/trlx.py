"""Public entry point: :func:`train` (parity: ``trlx/trlx.py:15-143``)."""
from __future__ import annotations

import os
import warnings
from typing import Callable, Dict, Iterable, List, Optional, Tuple

from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ilql_config, default_ppo_config, default_sft_config
from trlx_b200.utils import logging, set_seed
from trlx_b200.utils.loading import get_pipeline, get_trainer

logger = logging.get_logger(__name__)


def train(  # noqa: C901
    model_path: Optional[str] = None,
    reward_fn: Optional[Callable[[List[str], List[str], List[str]], List[float]]] = None,
    dataset: Optional[Iterable[Tuple[str, float]]] = None,
    samples: Optional[List[str]] = None,
    rewards: Optional[List[float]] = None,
    prompts: Optional[List[str]] = None,
    eval_prompts: Optional[List[str]] = None,
    metric_fn: Optional[Callable[[List[str], List[str], List[str]], Dict[str, List[float]]]] = None,
    config: Optional[TRLConfig] = None,
    stop_sequences: Optional[List[str]] = [],
):
    """Run online (``reward_fn`` + ``prompts``: PPO / RFT) or offline (``samples`` [+ ``rewards``]: ILQL / SFT) training.

    :param model_path: overrides ``config.model.model_path``
    :param reward_fn: ``reward_fn(samples, prompts, outputs, tokenizer=…, **metadata) -> List[float] | List[List[float]]``
        (a list per sample = dense per-token rewards)
    :param dataset: deprecated ``(samples, rewards)`` pair
    :param samples: strings, or lists of alternating (prompt, output, …) strings for dialogues
    :param rewards: one scalar per sample (→ ILQL); omit for SFT
    :param prompts: strings or dicts with a ``"prompt"`` key (other keys are forwarded to ``reward_fn``/``metric_fn``)
    :param eval_prompts: prompts for periodic evaluation
    :param metric_fn: ``metric_fn(samples, prompts, outputs, **metadata) -> Dict[str, List[float]]``
    :param config: :class:`TRLConfig`; a default is chosen from the arguments when omitted (deprecated)
    :param stop_sequences: generations are trimmed (and right-stripped) at the first occurrence of any of these
    :returns: the trainer
    """
    config = _resolve_config(config, model_path, online=bool(reward_fn), with_rewards=bool(rewards) or bool(dataset))
    set_seed(config.train.seed, config.train.parallel)
    if dataset:
        warnings.warn("the `dataset` argument is being depreciated, split it into `samples` and `rewards` instead")
        samples, rewards = dataset

    trainer_cls = get_trainer(config.train.trainer)
    trainer = trainer_cls(config=config, reward_fn=reward_fn, metric_fn=metric_fn, stop_sequences=stop_sequences,
                          **config.train.trainer_kwargs)

    global_batch = config.train.batch_size * int(os.environ.get("WORLD_SIZE", 1))
    new_tokens = config.method.gen_kwargs["max_new_tokens"]
    prompt_budget = config.train.seq_length - new_tokens
    if prompts is not None and prompt_budget <= 0:
        raise ValueError(f"train.seq_length ({config.train.seq_length}) leaves no room for prompts next to "
                         f"gen_kwargs.max_new_tokens ({new_tokens})")
    bos = trainer.tokenizer.bos_token

    def prompt_pipeline(items):
        return get_pipeline(config.train.pipeline)(items, prompt_budget, trainer.tokenizer,
                                                   add_special_tokens=config.model.model_arch_type == "seq2seq")

    if reward_fn:
        # online methods (PPO, RFT): roll out from the prompts; with none given, generation starts from BOS
        prompts = prompts or [bos] * global_batch
        eval_prompts = prompts[:global_batch] if eval_prompts is None else eval_prompts
        trainer.add_prompt_pipeline(prompt_pipeline(prompts))
    elif samples:
        # offline methods: ILQL when rewards come with the samples, SFT otherwise
        if rewards is None:
            trainer.make_experience(samples, config.train.seq_length)
        elif len(rewards) == len(samples):
            trainer.make_experience(samples, rewards, config.train.seq_length)
        else:
            raise ValueError(f"Number of samples {len(samples)} should match the number of rewards {len(rewards)}")
        eval_prompts = [bos] * global_batch if eval_prompts is None else eval_prompts
    else:
        raise ValueError("Either `samples` or `reward_fn` should be given for training")
    cap = int(os.environ.get("TRLX_B200_MAX_EVAL_PROMPTS", "0") or 0)
    if cap > 0 and len(eval_prompts) > cap:  # smoke runs on a CPU (scripts/smoke_examples.sh): keep evaluation short
        logger.info(f"TRLX_B200_MAX_EVAL_PROMPTS={cap}: evaluating on {cap} of {len(eval_prompts)} prompts")
        eval_prompts = eval_prompts[:cap]
    trainer.add_eval_pipeline(prompt_pipeline(eval_prompts))

    checkpoint = config.train.resume_from_checkpoint
    if checkpoint and os.path.exists(checkpoint):
        trainer.load(checkpoint)

    trainer.learn()
    if hasattr(trainer, "release_device_state"):
        trainer.release_device_state()  # captured CUDA graphs must not outlive the process group they reference
    return trainer


def _resolve_config(config: Optional[TRLConfig], model_path, online: bool, with_rewards: bool) -> TRLConfig:
    """The config the run uses: the caller's, or (deprecated) a default picked from which arguments were given; then the launch
    preset of ``python -m trlx_b200.launch`` (``TRLX_B200_PARALLEL``) and the ``model_path`` override."""
    if config is None:
        warnings.warn("Passing the `config` argument implicitly is depreciated, use or adapt some from "
                      "`trlx_b200/data/default_configs.py` instead")
        config = default_ppo_config() if online else (default_ilql_config() if with_rewards else default_sft_config())
    preset = os.environ.get("TRLX_B200_PARALLEL")
    if preset:
        import json

        config = config.evolve(train=dict(parallel=json.loads(preset)))
    if model_path:
        config.model.model_path = model_path
    return config
