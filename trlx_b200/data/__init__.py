from trlx_b200.data.types import BatchElement, GeneralElement, RLElement  # noqa: F401
