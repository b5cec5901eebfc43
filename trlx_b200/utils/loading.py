"""Name → class lookup for trainers and pipelines (parity: ``trlx/utils/loading.py``).  Importing this module
imports every trainer/pipeline module so their ``@register_*`` decorators run.  The Megatron-style
(``NeMo*Trainer``) names resolve to the in-repo tensor/pipeline-parallel trainers — unlike the reference snapshot,
whose NeMo backend cannot be imported."""
from __future__ import annotations

from typing import Callable

# registration side effects
from trlx_b200.pipeline import _DATAPIPELINE
from trlx_b200.pipeline.offline_pipeline import PromptPipeline  # noqa: F401
from trlx_b200.trainer import _TRAINERS, register_trainer  # noqa: F401
from trlx_b200.trainer.accelerate_ilql_trainer import AccelerateILQLTrainer  # noqa: F401
from trlx_b200.trainer.accelerate_ppo_trainer import AcceleratePPOTrainer  # noqa: F401
from trlx_b200.trainer.accelerate_rft_trainer import AccelerateRFTTrainer  # noqa: F401
from trlx_b200.trainer.accelerate_sft_trainer import AccelerateSFTTrainer  # noqa: F401
from trlx_b200.trainer.nemo_ilql_trainer import NeMoILQLTrainer  # noqa: F401
from trlx_b200.trainer.nemo_ppo_trainer import NeMoPPOTrainer  # noqa: F401
from trlx_b200.trainer.nemo_sft_trainer import NeMoSFTTrainer  # noqa: F401


def get_trainer(name: str) -> Callable:
    """Trainer class registered under ``name`` (case-insensitive)."""
    try:
        return _TRAINERS.get(name)
    except KeyError:
        raise Exception("Error: Trying to access a trainer that has not been registered") from None


def get_pipeline(name: str) -> Callable:
    """Pipeline class registered under ``name`` (case-insensitive)."""
    try:
        return _DATAPIPELINE.get(name)
    except KeyError:
        raise Exception("Error: Trying to access a pipeline that has not been registered") from None
