"""Small functional helpers with the reference's names and behaviour (reference src/util.py:5-58).
`Record` is the return type of `vAe` (src/model.py:73), so it is part of the drop-in surface."""
from collections.abc import Mapping
from functools import partial  # noqa: F401  (re-exported like the reference does)


def identity(x):
    return x


def comp(*fs):
    """right-to-left composition: comp(h, g, f)(x) == h(g(f(x)))  (src/util.py:10-13)."""
    if not fs:
        return identity

    def composed(x):
        for f in reversed(fs):
            x = f(x)
        return x
    return composed


class Record(Mapping):
    """attribute-style finite mapping (src/util.py:25-52): Record(a=1).a == Record(a=1)['a'] == 1;
    iteration yields keys; records merge left to right, keyword entries last."""

    def __init__(self, *records, **entries):
        for rec in records + (entries,):
            for k, v in rec.items():
                setattr(self, k, v)

    def __getitem__(self, key):
        try:
            return getattr(self, key)
        except AttributeError:
            raise KeyError(key)

    def __iter__(self):
        return iter(vars(self))

    def __len__(self):
        return len(vars(self))

    def __repr__(self):
        return repr(vars(self))


def select(record, *keys):
    """sub-record with only `keys` (src/util.py:55-57)."""
    return Record({k: record[k] for k in keys})
