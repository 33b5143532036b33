"""``TensorList`` — list of tensors with element-wise broadcasting arithmetic.

API-compatible with the reference container (``lib/tensorlist.py:5-180``): binary operators broadcast over a
second list or a scalar/tensor, unknown attributes resolve to ``torch.Tensor`` methods mapped over the elements,
integer indexing returns an element and slice/sequence indexing returns a ``TensorList``.  On this path it only
carries parameters and CG state between Python objects; the arithmetic of the optimiser itself runs inside
``libfrtm_b200`` (see ``model/optimizer.py``).  Dunder lookups are not forwarded, so ``torch.autograd`` and
``hasattr(x, '__torch_function__')`` see a plain list (the reference's version breaks on torch >= 2 there).
"""
from __future__ import annotations

import operator

import torch


def _is_seq(x):
    return isinstance(x, (TensorList, list))


class TensorList(list):
    def __init__(self, tensors=None):
        super().__init__(tensors if tensors is not None else [])

    # -- indexing --------------------------------------------------------------------------------------------
    def __getitem__(self, item):
        if isinstance(item, int):
            return list.__getitem__(self, item)
        if isinstance(item, (tuple, list)):
            return TensorList([list.__getitem__(self, i) for i in item])
        return TensorList(list.__getitem__(self, item))

    # -- arithmetic ------------------------------------------------------------------------------------------
    def _map2(self, other, fn, swap=False):
        if _is_seq(other):
            return TensorList([fn(b, a) if swap else fn(a, b) for a, b in zip(self, other)])
        return TensorList([fn(other, a) if swap else fn(a, other) for a in self])

    def _imap2(self, other, fn):
        if _is_seq(other):
            for i, b in enumerate(other):
                self[i] = fn(list.__getitem__(self, i), b)
        else:
            for i in range(len(self)):
                self[i] = fn(list.__getitem__(self, i), other)
        return self

    def __pos__(self):
        return TensorList([+e for e in self])

    def __neg__(self):
        return TensorList([-e for e in self])

    def __le__(self, other):
        return self._map2(other, operator.le)

    def __ge__(self, other):
        return self._map2(other, operator.ge)

    # -- list-flavoured helpers ------------------------------------------------------------------------------
    def concat(self, other):
        return TensorList(list.__add__(self, other))

    def copy(self):
        return TensorList(list.copy(self))

    def unroll(self):
        flat = TensorList()
        for t in self:
            if isinstance(t, TensorList):
                flat.extend(t.unroll())
            else:
                flat.append(t)
        return flat

    def list(self):
        return list(self)

    def attribute(self, attr: str, *args):
        return TensorList([getattr(e, attr, *args) for e in self])

    def apply(self, fn):
        return TensorList([fn(e) for e in self])

    def __getattr__(self, name):
        if name.startswith("__") or not hasattr(torch.Tensor, name):
            raise AttributeError("'TensorList' object has not attribute '{}'".format(name))

        def mapped(*args, **kwargs):
            return TensorList([getattr(e, name)(*args, **kwargs) for e in self])

        return mapped

    @staticmethod
    def _iterable(a):
        return _is_seq(a)


def _install():
    table = dict(add=operator.add, sub=operator.sub, mul=operator.mul, truediv=operator.truediv,
                 matmul=operator.matmul, mod=operator.mod)
    inplace = dict(add=operator.iadd, sub=operator.isub, mul=operator.imul, truediv=operator.itruediv,
                   matmul=operator.imatmul)
    for name, fn in table.items():
        setattr(TensorList, "__%s__" % name, (lambda f: lambda self, o: self._map2(o, f))(fn))
        setattr(TensorList, "__r%s__" % name, (lambda f: lambda self, o: self._map2(o, f, swap=True))(fn))
    for name, fn in inplace.items():
        setattr(TensorList, "__i%s__" % name, (lambda f: lambda self, o: self._imap2(o, f))(fn))


_install()


def tensor_operation(op):
    """Decorator: lets ``op`` accept TensorLists in its first one or two positional arguments (``:183-207``)."""
    import functools

    @functools.wraps(op)
    def wrapped(*args, **kwargs):
        if not args:
            raise ValueError("Must be at least one argument without keyword (i.e. operand).")
        a_list = isinstance(args[0], TensorList)
        b_list = len(args) > 1 and isinstance(args[1], TensorList)
        if a_list and b_list:
            return TensorList([op(a, b, *args[2:], **kwargs) for a, b in zip(args[0], args[1])])
        if a_list:
            return TensorList([op(a, *args[1:], **kwargs) for a in args[0]])
        if b_list:
            return TensorList([op(args[0], b, *args[2:], **kwargs) for b in args[1]])
        return op(*args, **kwargs)

    return wrapped
