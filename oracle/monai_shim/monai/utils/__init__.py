"""Shim of monai.utils.{ensure_tuple_rep, deprecated_arg}."""


def ensure_tuple_rep(tup, dim):
    if isinstance(tup, (list, tuple)):
        if len(tup) != dim:
            raise ValueError(f"sequence must have length {dim}, got {len(tup)}")
        return tuple(tup)
    return (tup,) * dim


def deprecated_arg(*_a, **_k):
    def deco(fn):
        return fn
    return deco
