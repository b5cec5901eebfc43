class _TT:
    def __class_getitem__(cls, item):
        return cls


TensorType = _TT
