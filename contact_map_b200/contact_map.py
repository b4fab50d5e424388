"""ContactFrequency / ContactDifference with the reference's Python API, computed on the GPU.

Mirror of /root/reference/contact_map/contact_map.py (ContactObject :109-615,
ContactFrequency :634-800, ContactDifference :803-986).  What changes underneath:

* ``ContactFrequency._build_contact_map`` (reference :717-744: a Python loop over frames
  around ``md.compute_neighborlist`` and set/Counter algebra) hands the whole trajectory to
  the sm_100a engine (``engine.ContactEngine``) and receives sparse count arrays;
* the ``Counter`` objects the reference keeps in ``_atom_contacts`` / ``_residue_contacts``
  are materialised lazily from those arrays (same keys: ``frozenset`` of REAL atom /
  residue indices; same values: integer frame counts).

There is no CPU implementation in this package: without a CUDA device the constructors
raise ``engine.EngineError``.
"""
import collections
import io
import json
import pickle
import warnings

import numpy as np

from .contact_count import ContactCount, PairTable
from .trajectory import topology_arrays

try:
    import torch
except ImportError:  # pragma: no cover - torch is part of the runtime image
    torch = None


def residue_neighborhood(residue, n=1):
    """Indices of the residues within ``n`` of ``residue`` in sequence AND in its chain
    (reference contact_map.py:41-60)."""
    in_chain = set(res.index for res in residue.chain.residues)
    return [idx for idx in range(residue.index - n, residue.index + n + 1) if idx in in_chain]


def _range_from_iterable(iterable):
    values = sorted(iterable)
    return (values[0], values[-1] + 1)


class ContactsDict(object):
    """``contacts['atom']`` / ``contacts['residue']`` access (reference :76-106)."""

    _ATOM = ("atom", "atoms")
    _RESIDUE = ("residue", "residues", "res")

    def __init__(self, contacts):
        self.contacts = contacts

    def __getitem__(self, atom_or_res):
        if atom_or_res in self._ATOM:
            return self.contacts.atom_contacts
        if atom_or_res in self._RESIDUE:
            return self.contacts.residue_contacts
        raise RuntimeError("Bad value for atom_or_res: " + str(atom_or_res))


class AtomIndexer(object):
    """Real <-> used ("sliced") atom index maps (reference atom_indexer.py:32-106).

    On the GPU atom slicing collapses into a gather index list (``all_atoms``); this object
    keeps the dictionaries the reference exposes for API compatibility.
    """

    def __init__(self, topology, real_query, real_haystack, all_atoms, sliced):
        self.all_atoms = all_atoms
        self.sliced = sliced
        if sliced:
            self.sliced_idx = {real: pos for pos, real in enumerate(all_atoms)}
        else:
            self.sliced_idx = {a: a for a in range(topology.n_atoms)}
        self.real_idx = {pos: real for real, pos in self.sliced_idx.items()}
        self.query = set(self.sliced_idx[q] for q in real_query)
        self.haystack = set(self.sliced_idx[h] for h in real_haystack)

    def convert_atom_contacts(self, atom_contacts):
        if not self.sliced:
            return atom_contacts
        return collections.Counter({frozenset(self.real_idx[a] for a in pair): value
                                    for pair, value in atom_contacts.items()})

    def slice_trajectory(self, trajectory):
        if self.sliced and len(self.all_atoms) < trajectory.topology.n_atoms:
            return trajectory.atom_slice(self.all_atoms)
        return trajectory


class ContactObject(object):
    """Shared set-up of all contact analyses (reference :109-615)."""

    # kept for API compatibility: slicing is a pure performance switch in the reference
    # (tests flip it globally); here the used atoms are always gathered.
    _class_use_atom_slice = None

    def __init__(self, topology, query, haystack, cutoff, n_neighbors_ignored):
        self._topology = topology
        if query is None or haystack is None:
            default = topology.select("not water and symbol != 'H'")
            query = default if query is None else query
            haystack = default if haystack is None else haystack
        self._cutoff = cutoff
        # sorted unique index arrays drive the GPU path; the reference's containers (sets of plain
        # ints, a sorted tuple) are built from them on first access (see the properties below)
        self._q_arr = _sorted_unique(query)
        self._h_arr = self._q_arr if haystack is query else _sorted_unique(haystack)
        self._all_arr = self._q_arr if self._q_arr is self._h_arr or np.array_equal(self._q_arr, self._h_arr) \
            else np.union1d(self._q_arr, self._h_arr)
        self._lazy = {}
        self._use_atom_slice = self._set_atom_slice(self._all_arr)
        if getattr(self, "_indexer", None) is None:
            self._indexer = None          # built on first access (two dicts over all atoms)
        self._n_neighbors_ignored = n_neighbors_ignored

    @property
    def indexer(self):
        """Real <-> used atom index maps (reference ``indexer`` attribute), built lazily."""
        if self._indexer is None:
            self._indexer = AtomIndexer(self._topology, self._query, self._haystack, self._all_atoms,
                                        self._use_atom_slice)
        return self._indexer

    @indexer.setter
    def indexer(self, value):
        self._indexer = value

    # -- the reference's containers (contact_map.py:132-139), built lazily from the arrays ------
    def _lazy_get(self, name, build):
        value = self._lazy.get(name)
        if value is None:
            value = self._lazy[name] = build()
        return value

    def _residues_of(self, atoms):
        atom_res, _ = topology_arrays(self._topology)
        return set(np.unique(atom_res[atoms]).tolist())

    @property
    def _query(self):
        return self._lazy_get("query", lambda: set(self._q_arr.tolist()))

    @property
    def _haystack(self):
        return self._lazy_get("haystack", lambda: set(self._h_arr.tolist()))

    @property
    def _query_res_idx(self):
        return self._lazy_get("query_res", lambda: self._residues_of(self._q_arr))

    @property
    def _haystack_res_idx(self):
        return self._lazy_get("haystack_res", lambda: self._residues_of(self._h_arr))

    @property
    def _all_atoms(self):
        return self._lazy_get("all_atoms", lambda: tuple(self._all_arr.tolist()))

    @property
    def _all_residues(self):
        return self._lazy_get("all_residues", lambda: self._residues_of(self._all_arr))

    def _set_atom_slice(self, all_atoms):
        if self._class_use_atom_slice is None:
            return len(all_atoms) < self._topology.n_atoms
        return self._class_use_atom_slice

    @classmethod
    def from_contacts(cls, atom_contacts, residue_contacts, topology, query=None, haystack=None,
                      cutoff=0.45, n_neighbors_ignored=2, indexer=None):
        obj = cls.__new__(cls)
        obj.indexer = indexer
        ContactObject.__init__(obj, topology, query, haystack, cutoff, n_neighbors_ignored)
        obj._set_counts(atom_contacts, residue_contacts)
        return obj

    # -- count storage: PairTable and/or Counter, converted on demand ------------------
    def _set_counts(self, atom_contacts, residue_contacts):
        self._atom_table = self._residue_table = None
        self._atom_counter = self._residue_counter = None
        for kind, value in (("atom", atom_contacts), ("residue", residue_contacts)):
            if isinstance(value, ContactCount):
                value = value.counter
            if isinstance(value, PairTable):
                setattr(self, "_%s_table" % kind, value)
            else:
                setattr(self, "_%s_counter" % kind, value)

    def _counter_of(self, kind):
        counter = getattr(self, "_%s_counter" % kind)
        if counter is None:
            counter = getattr(self, "_%s_table" % kind).to_counter()
            setattr(self, "_%s_counter" % kind, counter)
        return counter

    def _table_of(self, kind):
        # Note: a Counter handed out by ``_atom_contacts`` and then modified IN PLACE is not
        # seen here; assigning to ``_atom_contacts`` (what ``+=`` does) is.
        table = getattr(self, "_%s_table" % kind)
        if table is None:
            table = PairTable.from_counter(getattr(self, "_%s_counter" % kind))
            setattr(self, "_%s_table" % kind, table)
        return table

    @property
    def _atom_contacts(self):
        return self._counter_of("atom")

    @_atom_contacts.setter
    def _atom_contacts(self, value):
        self._atom_counter, self._atom_table = value, None

    @property
    def _residue_contacts(self):
        return self._counter_of("residue")

    @_residue_contacts.setter
    def _residue_contacts(self, value):
        self._residue_counter, self._residue_table = value, None

    # -- plain properties ----------------------------------------------------------------
    @property
    def contacts(self):
        return ContactsDict(self)

    @property
    def cutoff(self):
        return self._cutoff

    @property
    def n_neighbors_ignored(self):
        return self._n_neighbors_ignored

    @property
    def query(self):
        return list(self._query)

    @property
    def haystack(self):
        return list(self._haystack)

    @property
    def all_atoms(self):
        return list(self._all_atoms)

    @property
    def topology(self):
        return self._topology

    @property
    def use_atom_slice(self):
        return self._use_atom_slice

    @property
    def query_range(self):
        return _range_from_iterable(self._query)

    @property
    def haystack_range(self):
        return _range_from_iterable(self._haystack)

    @property
    def query_residue_range(self):
        return _range_from_iterable(self._query_res_idx)

    @property
    def haystack_residue_range(self):
        return _range_from_iterable(self._haystack_res_idx)

    @property
    def query_residues(self):
        return set(self.topology.atom(a).residue for a in self._query)

    @property
    def haystack_residues(self):
        return set(self.topology.atom(a).residue for a in self._haystack)

    @property
    def _residue_ignore_atom_idxs(self):
        """query residue index -> used-atom indices ignored for it (reference :442-461).
        Not used by the GPU path (the kernel evaluates the equivalent predicate); kept so the
        predicate can be inspected and tested."""
        used = {real: pos for pos, real in enumerate(self._all_atoms)} if self._use_atom_slice \
            else None
        result = {}
        atom_res, _ = topology_arrays(self.topology)
        for residue_idx in sorted(set(int(atom_res[q]) for q in self._query)):
            residue = self.topology.residue(residue_idx)
            ignored = set()
            for idx in residue_neighborhood(residue, self._n_neighbors_ignored):
                ignored.update(atom.index for atom in self.topology.residue(idx).atoms)
            if used is not None:
                ignored = set(used[a] for a in ignored if a in used)
            result[residue_idx] = ignored
        return result

    def __hash__(self):
        return hash((self.cutoff, self.n_neighbors_ignored, frozenset(self._query),
                     frozenset(self._haystack), self.topology))

    def __eq__(self, other):
        return (self.cutoff == other.cutoff
                and self.n_neighbors_ignored == other.n_neighbors_ignored
                and self.query == other.query and self.haystack == other.haystack
                and self.topology == other.topology)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __sub__(self, other):
        return ContactDifference(positive=self, negative=other)

    def _check_compatibility(self, other, err=AssertionError):
        failed = {}
        lines = []
        for attr in ("cutoff", "topology", "query", "haystack", "n_neighbors_ignored"):
            mine, theirs = getattr(self, attr), getattr(other, attr)
            if mine != theirs:
                failed[attr] = (mine, theirs)
                lines.append("        %s: %s != %s\n" % (attr, mine, theirs))
        if failed and err is not None:
            raise err("Incompatible ContactObjects:\n" + "".join(lines))
        return failed

    # -- persistence (same formats as the reference :202-394; host-side only) --------------
    def save_to_file(self, filename, mode="w"):
        with open(filename, mode + "b") as f:
            pickle.dump(self, f)

    @classmethod
    def from_file(cls, filename):
        with open(filename, "rb") as f:
            return pickle.load(f)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_atom_counter"] = self._counter_of("atom") if self._has_counts() else None
        state["_residue_counter"] = self._counter_of("residue") if self._has_counts() else None
        state["_atom_table"] = state["_residue_table"] = None
        return state

    def _has_counts(self):
        return (getattr(self, "_atom_counter", None) is not None
                or getattr(self, "_atom_table", None) is not None)

    # -- dict / JSON round trip: the reference's format (contact_map.py:202-333), so that files
    #    written by either package load in the other --------------------------------------------
    def to_dict(self):
        return {
            "topology": self._serialize_topology(self.topology),
            "cutoff": self._cutoff,
            "query": [int(v) for v in self._query],
            "haystack": [int(v) for v in self._haystack],
            "query_res_idx": [int(v) for v in self._query_res_idx],
            "haystack_res_idx": [int(v) for v in self._haystack_res_idx],
            "all_atoms": tuple(int(v) for v in self._all_atoms),
            "all_residues": tuple(int(v) for v in self._all_residues),
            "n_neighbors_ignored": self._n_neighbors_ignored,
            "atom_contacts": self._serialize_contact_counter(self._atom_contacts),
            "residue_contacts": self._serialize_contact_counter(self._residue_contacts),
            "use_atom_slice": self._use_atom_slice,
        }

    @classmethod
    def from_dict(cls, dct, topology=None):
        """Rebuild from ``to_dict`` output of this package OR of the reference.  ``topology``
        overrides the serialised one (kept for callers of the earlier two-argument form)."""
        if topology is None:
            topology = cls._deserialize_topology(dct["topology"])
        obj = cls.__new__(cls)
        obj.indexer = None
        ContactObject.__init__(obj, topology, [int(v) for v in dct["query"]],
                               [int(v) for v in dct["haystack"]], dct["cutoff"],
                               dct["n_neighbors_ignored"])
        if dct.get("use_atom_slice") is not None:
            obj._use_atom_slice = dct["use_atom_slice"]
        obj._set_counts(cls._deserialize_contact_counter(dct["atom_contacts"]),
                        cls._deserialize_contact_counter(dct["residue_contacts"]))
        if "n_frames" in dct:
            obj._n_frames = dct["n_frames"]
        return obj

    def to_json(self):
        return json.dumps(self.to_dict())

    @classmethod
    def from_json(cls, json_string, topology=None):
        return cls.from_dict(json.loads(json_string), topology)

    @staticmethod
    def _serialize_topology(topology):
        """JSON of (atom table, bonds) -- reference contact_map.py:278-283."""
        table, bonds = topology.to_dataframe()
        return json.dumps((table.to_json(), np.asarray(bonds).tolist()))

    @staticmethod
    def _deserialize_topology(topology_json):
        """Reference contact_map.py:269-276, into this package's array-backed Topology."""
        import pandas as pd
        from .trajectory import Topology
        table, bonds = json.loads(topology_json)
        return Topology.from_dataframe(pd.read_json(io.StringIO(table)), np.array(bonds))

    @staticmethod
    def _serialize_contact_counter(counter):
        return json.dumps({json.dumps(sorted(int(v) for v in key)): value
                           for key, value in counter.items()})

    @staticmethod
    def _deserialize_contact_counter(json_string):
        return collections.Counter({frozenset(json.loads(key)): value
                                    for key, value in json.loads(json_string).items()})

    @property
    def atom_contacts(self):
        raise NotImplementedError()

    @property
    def residue_contacts(self):
        raise NotImplementedError()

    # -- most-common helpers (reference :493-554), pure host work on finished counters ------
    def most_common_atoms_for_residue(self, residue):
        residue = residue if hasattr(residue, "index") else self.topology.residue(residue)
        members = set(atom.index for atom in residue.atoms)
        return [([self.topology.atom(a) for a in pair], value)
                for pair, value in self.atom_contacts.most_common_idx() if pair & members]

    def most_common_atoms_for_contact(self, contact_pair):
        first, second = [r if hasattr(r, "index") else self.topology.residue(r)
                         for r in list(contact_pair)]
        atoms_1 = set(atom.index for atom in first.atoms)
        atoms_2 = set(atom.index for atom in second.atoms)
        wanted = set(frozenset((a, b)) for a in atoms_1 for b in atoms_2)
        return [([self.topology.atom(a) for a in pair], value)
                for pair, value in self.atom_contacts.most_common_idx() if frozenset(pair) in wanted]


def _sorted_unique(indices):
    """Sorted unique int64 array of an iterable of atom indices (already-sorted input is kept)."""
    arr = np.asarray(indices if isinstance(indices, np.ndarray) else list(indices), dtype=np.int64).reshape(-1)
    if arr.shape[0] > 1 and not bool(np.all(arr[1:] > arr[:-1])):
        arr = np.unique(arr)
    return arr


def _split_trajectory(trajectory):
    """(xyz, unit-cell vectors or None) of an mdtraj-like trajectory object."""
    xyz = trajectory.xyz
    have_cell = getattr(trajectory, "_have_unitcell", None)
    box = trajectory.unitcell_vectors if (have_cell is None or have_cell) else None
    return xyz, box


def _engine_system(contact_object, engine):
    """Upload the per-used-atom arrays of ``contact_object`` (cmb_set_system)."""
    atom_res, atom_chain = topology_arrays(contact_object.topology)
    used = contact_object._all_arr
    role = np.zeros(used.shape[0], dtype=np.uint8)
    role[np.searchsorted(used, contact_object._q_arr)] |= 1        # both are subsets of the sorted `used`
    role[np.searchsorted(used, contact_object._h_arr)] |= 2
    if contact_object.n_neighbors_ignored is None or contact_object.cutoff is None:
        raise ValueError("cutoff and n_neighbors_ignored must be set to compute contacts")
    if contact_object.n_neighbors_ignored < 0:
        raise ValueError("n_neighbors_ignored must be >= 0")
    engine.set_system(atom_res[used], atom_chain[used], role, contact_object.cutoff,
                      contact_object.n_neighbors_ignored)
    return used


def _used_coordinates(xyz, box, used, n_total, engine, keep_full_host=False):
    """Slice the coordinates to the used atoms (reference atom_indexer._atom_slice) and say
    whether they live on the device.  ``keep_full_host``: leave a host array that still holds all
    atoms untouched (the caller streams it through ``accumulate_host_gather``)."""
    on_device = torch is not None and isinstance(xyz, torch.Tensor) and xyz.is_cuda
    if on_device:
        xyz = xyz.contiguous()
        if used.shape[0] < n_total:
            xyz = engine.gather_atoms(xyz, used)
        if box is not None and not (isinstance(box, torch.Tensor) and box.is_cuda):
            box = torch.as_tensor(np.ascontiguousarray(box, dtype=np.float32), device=engine.device)
        if box is not None:
            box = box.reshape(-1, 3, 3).contiguous()
        return xyz, box, True
    if torch is not None and isinstance(xyz, torch.Tensor):
        xyz = xyz.numpy()
    xyz = np.asarray(xyz)
    if xyz.dtype != np.float32:
        xyz = xyz.astype(np.float32)
    if used.shape[0] < n_total and not (keep_full_host and xyz.flags["C_CONTIGUOUS"]):
        xyz = np.ascontiguousarray(xyz[:, used])
    elif not xyz.flags["C_CONTIGUOUS"]:
        xyz = np.ascontiguousarray(xyz)
    if box is not None:
        if torch is not None and isinstance(box, torch.Tensor):
            box = box.cpu().numpy()
        box = np.ascontiguousarray(box, dtype=np.float32).reshape(-1, 3, 3)
    return xyz, box, False


def _table_kind(table):
    """'hashed' or 'dense': which count-table layout a build used (diagnostics / tests)."""
    return "hashed" if hasattr(table, "slots") else "dense"


def _tables_to_pairs(engine, atom_table, res_table, used):
    """Sparse export of the device tables -> (atom PairTable, residue PairTable) in REAL indices."""
    views = engine.export_pair_views(atom_table, res_table) if hasattr(engine, "export_pair_views") else None
    if views is None:
        return _exports_to_pairs(engine, engine.export_many([atom_table, res_table]), used)
    # (row, column, count) triples split on the device; the index maps below copy out of the
    # recycled staging buffers
    (lo, hi, cnt), (rlo, rhi, rcnt) = views
    rmap = engine.residue_map
    # (int32 wherever the indices allow: PairTable widens on first access)
    small = used.shape[0] == 0 or int(used[-1]) < 2 ** 31
    umap = used.astype(np.int32) if small else used
    rmap = rmap.astype(np.int32) if rmap.shape[0] == 0 or int(rmap.max()) < 2 ** 31 else rmap.astype(np.int64)
    return (PairTable(umap[lo], umap[hi], np.array(cnt).view(np.int32)),
            PairTable(rmap[rlo], rmap[rhi], np.array(rcnt).view(np.int32)))


def _exports_to_pairs(engine, exports, used):
    """(flat index, count) exports of the atom and residue tables -> PairTables in REAL indices."""
    n = engine.n_used
    (idx, cnt), (ridx, rcnt) = exports
    lo, hi = np.divmod(idx, n)
    atoms = PairTable(used[lo], used[hi], cnt.astype(np.int64))
    rc = engine.n_res_c
    rmap = engine.residue_map.astype(np.int64)
    rlo, rhi = np.divmod(ridx, rc)
    residues = PairTable(rmap[rlo], rmap[rhi], rcnt.astype(np.int64))
    return atoms, residues


class ContactFrequency(ContactObject):
    """Contact frequency (atomic and residue) for a trajectory.

    Same signature and results as the reference class (contact_map.py:634-800).
    ``trajectory`` is any mdtraj-like object (see ``trajectory.py``); its ``xyz`` may be a
    numpy array (host; staged to the GPU in overlapped chunks) or a CUDA ``torch.Tensor``.
    """
    _class_use_atom_slice = None
    _pending_dep_msg = (
        "ContactFrequency will be renamed to ContactMap in version 0.8. "
        "Invoking it as ContactFrequency will be deprecated in 0.8. For "
        "more, see https://github.com/dwhswenson/contact_map/issues/82"
    )

    def __init__(self, trajectory, query=None, haystack=None, cutoff=0.45, n_neighbors_ignored=2):
        warnings.warn(self._pending_dep_msg, PendingDeprecationWarning)
        self._n_frames = len(trajectory)
        super(ContactFrequency, self).__init__(trajectory.topology, query, haystack, cutoff,
                                               n_neighbors_ignored)
        self._set_counts(*self._build_contact_map(trajectory))

    @classmethod
    def from_contacts(cls, atom_contacts, residue_contacts, n_frames, topology, query=None,
                      haystack=None, cutoff=0.45, n_neighbors_ignored=2, indexer=None):
        warnings.warn(cls._pending_dep_msg, PendingDeprecationWarning)
        obj = super(ContactFrequency, cls).from_contacts(atom_contacts, residue_contacts, topology,
                                                         query, haystack, cutoff,
                                                         n_neighbors_ignored, indexer)
        obj._n_frames = n_frames
        return obj

    @classmethod
    def from_file(cls, filename):
        warnings.warn(cls._pending_dep_msg, PendingDeprecationWarning)
        return super(ContactFrequency, cls).from_file(filename)

    def _build_contact_map(self, trajectory):
        """GPU replacement of the reference's frame loop (contact_map.py:717-744).
        Returns (atom PairTable, residue PairTable) of integer frame counts, REAL indices."""
        from .engine import get_engine
        engine = get_engine()
        used = _engine_system(self, engine)
        xyz, box = _split_trajectory(trajectory)
        xyz, box, on_device = _used_coordinates(xyz, box, used, self.topology.n_atoms, engine,
                                                keep_full_host=True)
        atom_table, res_table = engine.new_tables()
        if xyz.shape[0] > 0:
            if on_device:
                engine.accumulate(xyz, box, atom_table, res_table)
            elif xyz.shape[1] != used.shape[0]:     # host array with all atoms: fused gather
                engine.accumulate_host_gather(xyz, used, box, atom_table, res_table)
            else:
                engine.accumulate_host(xyz, box, atom_table, res_table)
        # the sparse export is the result: the device tables (29 MB at 2.7k atoms, 40 GB at 100k)
        # are dropped here, only their kind is remembered
        self._table_kind = _table_kind(atom_table)
        return _tables_to_pairs(engine, atom_table, res_table, used)

    def __hash__(self):
        return hash((super(ContactFrequency, self).__hash__(),
                     tuple(self._atom_contacts.items()), tuple(self._residue_contacts.items()),
                     self.n_frames))

    def __eq__(self, other):
        return (super(ContactFrequency, self).__eq__(other)
                and self._atom_contacts == other._atom_contacts
                and self._residue_contacts == other._residue_contacts
                and self.n_frames == other.n_frames)

    @property
    def n_frames(self):
        return self._n_frames

    def _combine(self, other, sign):
        self._check_compatibility(other)
        atoms = self._table_of("atom").combine(other._table_of("atom"), sign)
        residues = self._table_of("residue").combine(other._table_of("residue"), sign)
        self._set_counts(atoms, residues)
        self._n_frames += sign * other._n_frames

    def add_contact_frequency(self, other):
        """Add the counts of ``other`` (the reduce operator, reference :751-763)."""
        self._combine(other, +1)

    def subtract_contact_frequency(self, other):
        """Remove the counts of ``other`` (a sub-trajectory of this one; reference :765-782)."""
        self._combine(other, -1)

    @property
    def atom_contacts(self):
        """Atom pairs mapped to the fraction of frames with that contact."""
        return ContactCount(self._table_of("atom").scaled(self.n_frames), self.topology.atom,
                            self.query_range, self.haystack_range, self.topology.n_atoms)

    @property
    def residue_contacts(self):
        """Residue pairs mapped to the fraction of frames with that contact."""
        return ContactCount(self._table_of("residue").scaled(self.n_frames), self.topology.residue,
                            self.query_residue_range, self.haystack_residue_range,
                            self.topology.n_residues)

    def to_dict(self):
        """Reference contact_map.py:712-715: the ContactObject keys plus ``n_frames``."""
        dct = super(ContactFrequency, self).to_dict()
        dct["n_frames"] = self.n_frames
        return dct


CONTACT_MAP_ERROR = (
    "The ContactMap class has been removed. Please use ContactFrequency."
    " For more, see: https://github.com/dwhswenson/contact_map/issues/82"
)


def ContactMap(*args, **kwargs):
    raise RuntimeError(CONTACT_MAP_ERROR)


# ---------------------------------------------------------------------------------------
# ContactDifference
# ---------------------------------------------------------------------------------------
def _atoms_match(one, other):
    """Atoms equal up to residue naming (reference topology.py:16-23)."""
    elem = lambda a: getattr(a, "element", None) or getattr(a, "element_symbol", None)   # noqa: E731
    return (one.name.rsplit("-", 1)[-1] == other.name.rsplit("-", 1)[-1]
            and elem(one) == elem(other) and one.index == other.index
            and one.serial == other.serial)


def _residues_match(one, other):
    return (one.index == other.index and one.resSeq == other.resSeq
            and getattr(one, "segment_id", "") == getattr(other, "segment_id", ""))


def _reconcile_topologies(map0, map1, override_topology):
    """(atoms_ok, residues_ok, topology) for two maps with different topologies
    (reference topology.py:101-131)."""
    top0, top1 = map0.topology, map1.topology
    atoms = set(map0._all_atoms) & set(map1._all_atoms)
    residues = set(map0._all_residues) & set(map1._all_residues)
    if override_topology is not None and override_topology is not True and override_topology is not False:
        topology = top0 = top1 = override_topology
    elif top0.n_atoms >= top1.n_atoms:
        topology = top0.copy()
    else:
        topology = top1.copy()
    try:
        atoms_ok = all(_atoms_match(top0.atom(i), top1.atom(i)) for i in atoms)
    except IndexError:
        atoms_ok = False
    try:
        pairs = [(top0.residue(i), top1.residue(i)) for i in residues]
        residues_ok = bool(pairs)
    except IndexError:
        pairs, residues_ok = [], False
    renamed = []
    for res0, res1 in pairs:
        if not _residues_match(res0, res1):
            residues_ok = False
        elif res0.name != res1.name:
            renamed.append((res0.index, res0.name, res1.name))
    if renamed:
        if override_topology:
            names = getattr(topology, "residue_names", None)
            for idx, name0, name1 in renamed:
                if names is not None:
                    names[idx] = "/".join([name0, name1])
                else:
                    topology.residue(idx).name = "/".join([name0, name1])
        else:
            residues_ok = False
    return atoms_ok, residues_ok, topology


class ContactDifference(ContactObject):
    """Difference of two contact maps, ``positive - negative`` (reference :803-914)."""
    _allow_mismatched_atoms = False
    _allow_mismatched_residues = False
    _override_topology = True

    def __init__(self, positive, negative):
        self.positive = positive
        self.negative = negative
        failed = positive._check_compatibility(negative, err=None)
        params = {"topology": positive.topology, "query": positive.query,
                  "haystack": positive.haystack, "cutoff": positive.cutoff,
                  "n_neighbors_ignored": positive.n_neighbors_ignored}
        for attr in failed:
            if attr in ("query", "haystack"):
                params[attr] = set(getattr(positive, attr)) & set(getattr(negative, attr))
            elif attr in ("cutoff", "n_neighbors_ignored"):
                params[attr] = None
            else:
                params[attr] = self._fix_topology(positive, negative)
        self._all_atoms_intersect = set(positive._all_atoms) & set(negative._all_atoms)
        self._all_residues_intersect = set(positive._all_residues) & set(negative._all_residues)
        super(ContactDifference, self).__init__(params["topology"], params["query"],
                                                params["haystack"], params["cutoff"],
                                                params["n_neighbors_ignored"])

    def _fix_topology(self, positive, negative):
        atoms_ok, residues_ok, topology = _reconcile_topologies(positive, negative,
                                                                self._override_topology)
        hint = ("`diff = OverrideTopologyContactDifference(map1, map2, topology)` with a "
                "topology that contains all atoms and residues.")
        if not atoms_ok and not residues_ok:
            if not (self._allow_mismatched_atoms and self._allow_mismatched_residues):
                raise RuntimeError("The two different contact maps had atoms and residues that "
                                   "were not equal between the two topologies. If you want to "
                                   "compare these, use " + hint)
        if not atoms_ok and not self._allow_mismatched_atoms:
            raise RuntimeError("The two different contact maps had atoms that were not equal "
                               "between the two topologies. If you want to compare these use "
                               "`diff = AtomMismatchedContactDifference(map1, map2)`.\n"
                               "Alternatively, use " + hint)
        if not residues_ok and not self._allow_mismatched_residues:
            raise RuntimeError("The two different contact maps had residues that were not equal "
                               "between the two topologies. If you want to compare these use "
                               "`diff = ResidueMismatchedContactDifference(map1, map2)`.\n"
                               "Alternatively, use " + hint)
        return topology

    def __sub__(self, other):
        raise NotImplementedError

    @classmethod
    def from_contacts(cls, *args, **kwargs):
        raise NotImplementedError

    def to_dict(self):
        """Reference contact_map.py:838-851."""
        return {"positive": self.positive.to_json(), "negative": self.negative.to_json(),
                "positive_cls": self.positive.__class__.__name__,
                "negative_cls": self.negative.__class__.__name__}

    @classmethod
    def from_dict(cls, dct):
        """Reference contact_map.py:853-880."""
        supported = {"ContactFrequency": ContactFrequency}

        def rebuild(which):
            name = dct[which + "_cls"]
            if name not in supported:
                raise RuntimeError("Can't rebuild class " + name)
            return supported[name].from_json(dct[which])
        return cls(rebuild("positive"), rebuild("negative"))

    @staticmethod
    def _filtered_difference(pos_count, neg_count, selection, object_f, max_size):
        """freq_pos - freq_neg over the union of keys restricted to ``selection``; keys whose
        difference is exactly 0.0 stay (reference :907-914, ``Counter.subtract``)."""
        sel = np.fromiter((int(s) for s in selection), dtype=np.int64)
        pos = pos_count.table.restricted(sel)
        neg = neg_count.table.restricted(sel)
        base = int(max(pos.j.max(initial=0), neg.j.max(initial=0))) + 1
        keys = np.concatenate([pos.i * base + pos.j, neg.i * base + neg.j])
        vals = np.concatenate([pos.value.astype(np.float64), -neg.value.astype(np.float64)])
        uniq, inv = np.unique(keys, return_inverse=True)
        diff = np.zeros(uniq.shape[0], dtype=np.float64)
        # each key occurs at most once per side, so this is exactly one subtraction per key
        np.add.at(diff, inv, vals)
        return ContactCount(PairTable(uniq // base, uniq % base, diff), object_f, max_size=max_size)

    @property
    def atom_contacts(self):
        return self._filtered_difference(self.positive.atom_contacts, self.negative.atom_contacts,
                                         self._all_atoms_intersect, self.topology.atom,
                                         self.topology.n_atoms)

    @property
    def residue_contacts(self):
        return self._filtered_difference(self.positive.residue_contacts,
                                         self.negative.residue_contacts,
                                         self._all_residues_intersect, self.topology.residue,
                                         self.topology.n_residues)


class AtomMismatchedContactDifference(ContactDifference):
    """Contact map comparison when only the residues can be compared."""
    _allow_mismatched_atoms = True

    def _missing(self, *args, **kwargs):
        raise RuntimeError("Different atom indices involved between the two maps, so this does "
                           "not make sense.")

    most_common_atoms_for_contact = most_common_atoms_for_residue = _missing
    atom_contacts = property(_missing)
    haystack_residues = property(_missing)
    query_residues = property(_missing)


class ResidueMismatchedContactDifference(ContactDifference):
    """Contact map comparison when only the atoms can be compared."""
    _allow_mismatched_residues = True

    def _missing(self, *args, **kwargs):
        raise RuntimeError("Different residue indices involved between the two maps, so this "
                           "does not make sense.")

    most_common_atoms_for_contact = most_common_atoms_for_residue = _missing
    residue_contacts = property(_missing)
    _residue_ignore_atom_idxs = property(_missing)
    haystack_residues = property(_missing)
    query_residues = property(_missing)


class OverrideTopologyContactDifference(ContactDifference):
    """Contact map comparison with a user-provided topology."""

    def __init__(self, positive, negative, topology):
        self._override_topology = topology
        super(OverrideTopologyContactDifference, self).__init__(positive, negative)
