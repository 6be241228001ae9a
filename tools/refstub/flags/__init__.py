"""Minimal stand-in for the `py-flags` package (only what the reference's
rendering/__init__.py:11-17 and rendering/ray_caster.py:13-20 touch at import)."""


class _Member(int):
    def __new__(cls, value, name):
        obj = super().__new__(cls, value)
        obj._name = name
        return obj

    def to_simple_str(self):
        return self._name

    def __or__(self, other):
        return _Member(int(self) | int(other), "|".join((self._name, getattr(other, "_name", str(other)))))

    __ror__ = __or__


class _Meta(type):
    def __new__(mcs, name, bases, ns):
        members = []
        bit = 1
        for key, val in list(ns.items()):
            if not key.startswith("_") and val == ():
                m = _Member(bit, key)
                ns[key] = m
                members.append(m)
                bit <<= 1
        ns["_members"] = members
        ns["no_flags"] = _Member(0, "no_flags")
        return super().__new__(mcs, name, bases, ns)

    def __iter__(cls):
        return iter(cls._members)


class Flags(metaclass=_Meta):
    pass
