"""Public entry point: :func:`train` (parity: ``trlx/trlx.py:15-143``)."""
from __future__ import annotations

import os
import warnings
from typing import Callable, Dict, Iterable, List, Optional, Tuple

from trlx_b200.data.configs import TRLConfig
from trlx_b200.data.default_configs import default_ilql_config, default_ppo_config, default_sft_config
from trlx_b200.utils import set_seed
from trlx_b200.utils.loading import get_pipeline, get_trainer


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
    if config is None:
        warnings.warn("Passing the `config` argument implicitly is depreciated, use or adapt some from "
                      "`trlx_b200/data/default_configs.py` instead")
        if reward_fn:
            config = default_ppo_config()
        elif rewards:
            config = default_ilql_config()
        else:
            config = default_sft_config()

    if os.environ.get("TRLX_B200_PARALLEL"):  # launch preset (python -m trlx_b200.launch --config_file …)
        import json

        config = config.evolve(train=dict(parallel=json.loads(os.environ["TRLX_B200_PARALLEL"])))
    set_seed(config.train.seed, config.train.parallel)

    if dataset:
        warnings.warn("the `dataset` argument is being depreciated, split it into `samples` and `rewards` instead")
        samples, rewards = dataset
    if model_path:
        config.model.model_path = model_path

    trainer = get_trainer(config.train.trainer)(config=config, reward_fn=reward_fn, metric_fn=metric_fn,
                                                stop_sequences=stop_sequences, **config.train.trainer_kwargs)

    batch_size = config.train.batch_size * int(os.environ.get("WORLD_SIZE", 1))
    max_prompt_length = config.train.seq_length - config.method.gen_kwargs["max_new_tokens"]
    if prompts is not None and max_prompt_length <= 0:
        raise ValueError(f"train.seq_length ({config.train.seq_length}) leaves no room for prompts next to "
                         f"gen_kwargs.max_new_tokens ({config.method.gen_kwargs['max_new_tokens']})")
    seq2seq = config.model.model_arch_type == "seq2seq"
    pipeline_cls = get_pipeline(config.train.pipeline)

    if reward_fn:  # online
        prompts = prompts or [trainer.tokenizer.bos_token] * batch_size
        if eval_prompts is None:
            eval_prompts = prompts[:batch_size]
        trainer.add_prompt_pipeline(pipeline_cls(prompts, max_prompt_length, trainer.tokenizer, add_special_tokens=seq2seq))
    elif samples:  # offline
        if rewards is not None and len(samples) != len(rewards):
            raise ValueError(f"Number of samples {len(samples)} should match the number of rewards {len(rewards)}")
        if eval_prompts is None:
            eval_prompts = [trainer.tokenizer.bos_token] * batch_size
        if rewards is not None:
            trainer.make_experience(samples, rewards, config.train.seq_length)
        else:
            trainer.make_experience(samples, config.train.seq_length)
    else:
        raise ValueError("Either `samples` or `reward_fn` should be given for training")

    trainer.add_eval_pipeline(pipeline_cls(eval_prompts, max_prompt_length, trainer.tokenizer, add_special_tokens=seq2seq))

    if config.train.resume_from_checkpoint and os.path.exists(config.train.resume_from_checkpoint):
        trainer.load(config.train.resume_from_checkpoint)

    trainer.learn()
    if hasattr(trainer, "release_device_state"):
        trainer.release_device_state()  # captured CUDA graphs must not outlive the process group they reference
    return trainer
