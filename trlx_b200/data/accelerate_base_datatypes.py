from trlx_b200.data.types import (  # noqa: F401
    AccelerateRLBatchElement,
    AccelerateRLElement,
    PromptBatch,
    PromptElement,
)
