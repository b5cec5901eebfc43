from trlx_b200.data.types import (  # noqa: F401
    ILQLBatch,
    ILQLElement,
    ILQLSeq2SeqBatch,
    ILQLSeq2SeqElement,
    flatten_dataclass,
    unflatten_dataclass,
)
