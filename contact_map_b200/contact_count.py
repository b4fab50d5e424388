"""Result containers of the contact-map path.

``ContactCount`` mirrors the reference's return type
(/root/reference/contact_map/contact_count.py:75-129,332-402: ``counter``, ``n_x``,
``n_y``, ``max_size``, ``total_range``, ``most_common``, ``filter``, ``sparse_matrix``,
``df``).  Unlike the reference it is array-backed: the GPU hands back sparse
``(i, j, value)`` arrays (``PairTable``) and the ``collections.Counter`` keyed by
``frozenset({i, j})`` is only materialised when ``.counter`` is first read.
Plotting / networkx export are presentation layers outside the hot path
 and are not provided.
"""
import collections

import numpy as np


class PairTable(object):
    """Sparse symmetric table over unordered index pairs: arrays ``i < j`` and ``value``.

    ``i`` / ``j`` are int64 and integer values int64 for every reader.  The GPU export hands over
    int32 arrays (3.9 M pairs on the 100 000-atom system: fresh int64 copies of them cost more host
    time than the export itself); they are kept as they come and widened on first access."""

    __slots__ = ("_i", "_j", "_value")

    def __init__(self, i, j, value):
        self._i = self._index_array(i)
        self._j = self._index_array(j)
        self._value = np.asarray(value)

    @staticmethod
    def _index_array(a):
        a = np.asarray(a)
        return a if a.dtype in (np.int32, np.int64) else a.astype(np.int64)

    @property
    def i(self):
        if self._i.dtype != np.int64:
            self._i = self._i.astype(np.int64)
        return self._i

    @property
    def j(self):
        if self._j.dtype != np.int64:
            self._j = self._j.astype(np.int64)
        return self._j

    @property
    def value(self):
        if self._value.dtype in (np.int32, np.uint32):
            self._value = self._value.astype(np.int64)
        return self._value

    @classmethod
    def empty(cls, dtype=np.int64):
        return cls(np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, dtype))

    @classmethod
    def from_counter(cls, counter):
        n = len(counter)
        i = np.empty(n, np.int64)
        j = np.empty(n, np.int64)
        vals = []
        for k, (key, val) in enumerate(counter.items()):
            a, b = tuple(key) if len(key) == 2 else (tuple(key)[0], tuple(key)[0])
            i[k], j[k] = (a, b) if a < b else (b, a)
            vals.append(val)
        value = np.asarray(vals) if n else np.empty(0, np.int64)
        return cls(i, j, value)

    def __len__(self):
        return int(self._i.shape[0])

    def to_counter(self):
        """``Counter{frozenset({i, j}): value}`` with plain Python ints / floats."""
        ii, jj, vv = self.i.tolist(), self.j.tolist(), self.value.tolist()
        return collections.Counter(dict(zip(map(frozenset, zip(ii, jj)), vv)))

    def _packed(self, base):
        return self.i * np.int64(base) + self.j

    def combine(self, other, sign=1):
        """self + sign*other on integer counts; non-positive entries are dropped, which is
        what ``Counter.__iadd__`` / ``__isub__`` do (contact_map.py:751-782)."""
        base = int(max(self.j.max(initial=0), other.j.max(initial=0))) + 1
        keys = np.concatenate([self._packed(base), other._packed(base)])
        vals = np.concatenate([self.value.astype(np.int64), sign * other.value.astype(np.int64)])
        uniq, inv = np.unique(keys, return_inverse=True)
        total = np.zeros(uniq.shape[0], dtype=np.int64)
        np.add.at(total, inv, vals)
        keep = total > 0
        uniq, total = uniq[keep], total[keep]
        return PairTable(uniq // base, uniq % base, total)

    def scaled(self, denominator):
        """float(count) / denominator, exactly the reference's Python arithmetic
        (contact_map.py:787-789): IEEE double division of exactly represented ints."""
        return PairTable(self.i, self.j, self.value.astype(np.float64) / float(denominator))

    def restricted(self, allowed):
        """Entries whose two indices are both in ``allowed`` (ContactCount.filter)."""
        allowed = np.fromiter(allowed, dtype=np.int64) if not isinstance(allowed, np.ndarray) \
            else allowed.astype(np.int64)
        keep = np.isin(self.i, allowed) & np.isin(self.j, allowed)
        return PairTable(self.i[keep], self.j[keep], self.value[keep])


class _ContactPlotRange(object):
    """(min, max) extent of one axis; compares equal to the int / tuple it was built from."""

    def __init__(self, n):
        self.n = n
        if isinstance(n, collections.abc.Iterable):
            self.min, self.max = n[0], n[1]
        else:
            self.min, self.max = 0, n

    @property
    def range_length(self):
        return self.max - self.min

    def __eq__(self, other):
        if isinstance(other, (int, tuple)):
            return self.n == other
        if isinstance(other, _ContactPlotRange):
            return (self.n, self.min, self.max) == (other.n, other.min, other.max)
        return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return "_ContactPlotRange(%r)" % (self.n,)


def _key_ranges(table):
    """((min low, max low + 1), (min high, max high + 1)) over the pair keys."""
    if len(table) == 0:
        return (0, 0), (0, 0)
    return ((int(table.i.min()), int(table.i.max()) + 1),
            (int(table.j.min()), int(table.j.max()) + 1))


def make_x_y_ranges(n_x, n_y, table):
    """Axis ranges (plot_utils.py:100-125): given values win; otherwise the shorter of the
    low-key / high-key extents becomes y."""
    if n_x is None and n_y is None:
        low, high = _key_ranges(table)
        n_x, n_y = (high, low) if low[1] - low[0] > high[1] - high[0] else (low, high)
    elif n_x is None or n_y is None:
        raise ValueError("Either both n_x and n_y need to be defined or neither.")
    if isinstance(n_x, _ContactPlotRange):
        n_x = n_x.n
    if isinstance(n_y, _ContactPlotRange):
        n_y = n_y.n
    return _ContactPlotRange(n_x), _ContactPlotRange(n_y)


class ContactCount(object):
    """Contacts of one kind (atom or residue) with their values.

    Parameters follow the reference (contact_count.py:113-123); ``counter`` may be a
    ``collections.Counter`` keyed by ``frozenset`` pairs or a :class:`PairTable`.
    """

    def __init__(self, counter, object_f, n_x=None, n_y=None, max_size=None):
        if isinstance(counter, PairTable):
            self._table = counter
            self._counter = None
        else:
            self._counter = counter
            self._table = PairTable.from_counter(counter)
        self._object_f = object_f
        if len(self._table):
            self.total_range = (int(min(self._table.i.min(), self._table.j.min())),
                                int(max(self._table.i.max(), self._table.j.max())) + 1)
        else:
            self.total_range = (0, 0)
        self.n_x, self.n_y = make_x_y_ranges(n_x, n_y, self._table)
        if max_size is None:
            self.max_size = max([self.total_range[-1], self.n_x.max, self.n_y.max])
        else:
            self.max_size = max_size

    @property
    def counter(self):
        """``collections.Counter``: frozenset index pair -> value (materialised lazily)."""
        if self._counter is None:
            self._counter = self._table.to_counter()
        return self._counter

    @property
    def table(self):
        """Array view of the same data (:class:`PairTable`)."""
        return self._table

    @property
    def sparse_matrix(self):
        """Symmetric ``scipy.sparse`` matrix (DOK, like the reference) of the values."""
        import scipy.sparse
        t = self._table
        size = self.max_size
        rows = np.concatenate([t.i, t.j])
        cols = np.concatenate([t.j, t.i])
        vals = np.concatenate([t.value, t.value]).astype(np.float64)
        return scipy.sparse.coo_matrix((vals, (rows, cols)), shape=(size, size)).todok()

    @property
    def df(self):
        """``pandas`` sparse DataFrame of the contact matrix (NaN where there is no contact)."""
        import pandas as pd
        index = list(range(self.max_size))
        frame = pd.DataFrame.sparse.from_spmatrix(self.sparse_matrix.tocoo(), index=index,
                                                  columns=index)
        return frame.astype(pd.SparseDtype("float", np.nan))

    def most_common_idx(self):
        """[(frozenset index pair, value)] in decreasing order of value."""
        return self.counter.most_common()

    def most_common(self, obj=None):
        """Like :meth:`most_common_idx` with topology objects as keys, optionally only the
        contacts that involve ``obj``."""
        wanted = None if obj is None else obj.index
        return [([self._object_f(idx) for idx in pair], value)
                for pair, value in self.most_common_idx()
                if wanted is None or wanted in pair]

    def filter(self, idx):
        """New ContactCount restricted to pairs whose members are both in ``idx``."""
        sub = self._table.restricted(np.fromiter((int(v) for v in idx), dtype=np.int64))
        return ContactCount(collections.Counter(sub.to_counter()), self._object_f, self.n_x, self.n_y)
