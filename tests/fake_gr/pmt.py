"""Minimal stand-in for GNU Radio's `pmt` module, enough to drive wifi_b200.gr_adapter in tests on
machines without GNU Radio.  PMTs are plain Python objects; only the calls the adapter makes exist."""
import numpy as np


class _Sym(str):
    pass


class _Pair(tuple):
    pass


def intern(s):
    return _Sym(s)


def cons(a, b):
    return _Pair((a, b))


def car(p):
    return p[0]


def cdr(p):
    return p[1]


def make_dict():
    return {}


def dict_add(d, k, v):
    n = dict(d)
    n[k] = v
    return n


def is_dict(x):
    return isinstance(x, dict)


def from_long(v):
    return int(v)


def from_double(v):
    return float(v)


def to_python(x):
    if isinstance(x, dict):
        return {str(k): v for k, v in x.items()}
    return x


def init_u8vector(n, vals):
    assert n == len(vals)
    return np.array(vals, dtype=np.uint8)


def u8vector_elements(v):
    return [int(b) for b in v]


def init_c32vector(n, vals):
    assert n == len(vals)
    return np.array(vals, dtype=np.complex64)
