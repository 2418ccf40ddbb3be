"""Result containers.

The reference returns ``monty.collections.AttrDict`` everywhere (part_decoder.py:22, object_decoder.py:19,
part_encoder.py:19); ``monty`` is not a dependency here, so an equivalent is shipped.  ``LazyAttrDict`` adds
on-demand entries: the fused kernels never materialise the B x K x H x W warped-template tensor, so entries such
as ``transformed_templates`` are computed by the render kernel only when somebody reads them.
"""


class AttrDict(dict):
    """dict with attribute access (``res.vote``, ``res.update(...)``, ``res.get(...)``)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name) from None


class _Thunk:
    __slots__ = ('fn',)

    def __init__(self, fn):
        self.fn = fn


class LazyAttrDict(AttrDict):
    """AttrDict whose values may be deferred: ``d.set_lazy(key, fn)`` stores ``fn`` and calls it on first read."""

    def set_lazy(self, key, fn):
        dict.__setitem__(self, key, _Thunk(fn))

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _Thunk):
            v = v.fn()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def is_materialized(self, key):
        return not isinstance(dict.__getitem__(self, key), _Thunk)

    # every way of reading the mapping resolves the deferred entries it touches, so that dict(res), {**res}, res.copy(),
    # res.pop(k), pickling or a collate function never see the internal thunks
    def keys(self):
        return list(dict.keys(self))

    def __iter__(self):
        return iter(self.keys())

    def pop(self, key, *default):
        if key in self:
            value = self[key]
            dict.__delitem__(self, key)
            return value
        if default:
            return default[0]
        raise KeyError(key)

    def popitem(self):
        key = next(reversed(dict.keys(self)))
        return key, self.pop(key)

    def setdefault(self, key, default=None):
        if key not in self:
            dict.__setitem__(self, key, default)
        return self[key]

    def copy(self):
        return AttrDict(self.items())

    def to_dict(self):
        """plain dict with every deferred entry computed"""
        return dict(self.items())

    def __reduce__(self):
        return (AttrDict, (self.to_dict(),))
