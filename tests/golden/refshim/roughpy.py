"""Stand-in for roughpy's Lie-basis enumeration (LinearNeuralCDEs.py:111-116): the Hall basis of the reference's
own data_dir/hall_set.py, keys printed as roughpy prints them ("3", "[1,2]", "[1,[1,2]]").  That roughpy 0.2.0
enumerates its Hall basis in this same order is an ASSUMPTION (equal by inspection at depth <= 2)."""
DPReal = "DPReal"


class _Key:
    def __init__(self, s):
        self.s = s

    def __str__(self):
        return self.s


class _Basis:
    def __init__(self, width, depth):
        from data_dir.hall_set import HallSet

        self.hs = HallSet(width, depth)

    def _key(self, k):
        l, r = self.hs.data[k]
        if l == 0:
            return str(r)
        return f"[{self._key(l)},{self._key(r)}]"

    def size(self, depth):
        return len(self.hs.data) - 1

    def index_to_key(self, i):
        return _Key(self._key(i + 1))


class _Ctx:
    def __init__(self, width, depth):
        self.lie_basis = _Basis(width, depth)


def get_context(width, depth, coeffs=None):
    return _Ctx(width, depth)
