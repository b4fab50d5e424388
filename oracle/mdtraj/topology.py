"""Object topology for the oracle's mdtraj stand-in (test infrastructure only).

Covers what the reference library and its tests touch: chain/residue/atom
objects with indices, ``select`` for a small subset of the MDTraj DSL, ``copy``,
``subset``, ``join``, data-frame round trip, atom insertion/removal, equality
and hashing, and the writable ``_atoms`` / ``_numAtoms`` attributes used by the
reference's ``_atom_slice`` shortcut (contact_map/atom_indexer.py:18-19).
"""
import re

import numpy as np

from . import element as elem

_WATER_RESIDUES = {"H2O", "HHO", "OHH", "HOH", "OH2", "SOL", "WAT", "TIP", "TIP2", "TIP3", "TIP4"}
_PROTEIN_RESIDUES = {
    "ACE", "NME", "ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU",
    "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL", "HID", "HIE", "HIP", "CYX",
    "ASH", "GLH", "LYN", "NH2", "NMA", "MSE",
}


class Atom(object):
    def __init__(self, name, element, index, residue, serial=None):
        self.name = name
        self.element = element
        self.index = index
        self.residue = residue
        self.serial = serial

    @property
    def segment_id(self):
        return self.residue.segment_id

    @property
    def n_bonds(self):
        return 0

    def __eq__(self, other):
        if not isinstance(other, Atom):
            return False
        return (self.index == other.index and self.name == other.name
                and self.element == other.element
                and self.residue.index == other.residue.index
                and self.residue.name == other.residue.name)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.index, self.name, self.residue.index, self.residue.name))

    def __str__(self):
        return "%s-%s" % (self.residue, self.name)

    __repr__ = __str__


class Residue(object):
    def __init__(self, name, index, chain, resSeq, segment_id=""):
        self.name = name
        self.index = index
        self.chain = chain
        self.resSeq = resSeq
        self.segment_id = segment_id
        self._atoms = []

    @property
    def atoms(self):
        return iter(self._atoms)

    def atom(self, index):
        return self._atoms[index]

    @property
    def n_atoms(self):
        return len(self._atoms)

    @property
    def is_water(self):
        return self.name in _WATER_RESIDUES

    @property
    def is_protein(self):
        return self.name in _PROTEIN_RESIDUES

    def __eq__(self, other):
        if not isinstance(other, Residue):
            return False
        return (self.index == other.index and self.name == other.name
                and self.resSeq == other.resSeq and self.chain.index == other.chain.index)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.index, self.name, self.resSeq, self.chain.index))

    def __str__(self):
        return "%s%s" % (self.name, self.resSeq)

    __repr__ = __str__


class Chain(object):
    def __init__(self, index, topology, chain_id=None):
        self.index = index
        self.topology = topology
        self.chain_id = chain_id
        self._residues = []

    @property
    def residues(self):
        return iter(self._residues)

    def residue(self, index):
        return self._residues[index]

    @property
    def n_residues(self):
        return len(self._residues)

    @property
    def atoms(self):
        for res in self._residues:
            for atom in res._atoms:
                yield atom

    @property
    def n_atoms(self):
        return sum(r.n_atoms for r in self._residues)

    def __eq__(self, other):
        return isinstance(other, Chain) and self.index == other.index

    def __hash__(self):
        return hash(("chain", self.index))


class Topology(object):
    def __init__(self):
        self._chains = []
        self._numResidues = 0
        self._numAtoms = 0
        self._bonds = []
        self._atoms = []
        self._residues = []

    # -- construction -------------------------------------------------
    def add_chain(self, chain_id=None):
        chain = Chain(len(self._chains), self, chain_id)
        self._chains.append(chain)
        return chain

    def add_residue(self, name, chain, resSeq=None, segment_id=""):
        if resSeq is None:
            resSeq = self._numResidues
        residue = Residue(name, self._numResidues, chain, resSeq, segment_id)
        self._residues.append(residue)
        self._numResidues += 1
        chain._residues.append(residue)
        return residue

    def add_atom(self, name, element, residue, serial=None):
        if element is None:
            element = elem.virtual
        atom = Atom(name, element, self._numAtoms, residue, serial=serial)
        self._atoms.append(atom)
        self._numAtoms += 1
        residue._atoms.append(atom)
        return atom

    def insert_atom(self, name, element, residue, index=None, rindex=None, serial=None):
        if index is None:
            index = self._numAtoms
        if rindex is None:
            rindex = len(residue._atoms)
        atom = Atom(name, element, index, residue, serial=serial)
        for a in self._atoms[index:]:
            a.index += 1
        self._atoms.insert(index, atom)
        self._numAtoms += 1
        residue._atoms.insert(rindex, atom)
        return atom

    def delete_atom_by_index(self, index):
        atom = self._atoms[index]
        atom.residue._atoms.remove(atom)
        del self._atoms[index]
        self._numAtoms -= 1
        for a in self._atoms[index:]:
            a.index -= 1
        self._bonds = [b for b in self._bonds if atom not in b]

    def add_bond(self, a1, a2, type=None, order=None):
        self._bonds.append((a1, a2))

    # -- access -------------------------------------------------------
    def atom(self, index):
        return self._atoms[index]

    def residue(self, index):
        return self._residues[index]

    def chain(self, index):
        return self._chains[index]

    @property
    def atoms(self):
        for chain in self._chains:
            for atom in chain.atoms:
                yield atom

    @property
    def residues(self):
        for chain in self._chains:
            for res in chain._residues:
                yield res

    @property
    def chains(self):
        return iter(self._chains)

    @property
    def bonds(self):
        return iter(self._bonds)

    @property
    def n_atoms(self):
        return len(self._atoms)

    @property
    def n_residues(self):
        return len(self._residues)

    @property
    def n_chains(self):
        return len(self._chains)

    @property
    def n_bonds(self):
        return len(self._bonds)

    # -- copies -------------------------------------------------------
    def copy(self):
        out = Topology()
        for chain in self._chains:
            c = out.add_chain(chain.chain_id)
            for res in chain._residues:
                r = out.add_residue(res.name, c, res.resSeq, res.segment_id)
                for atom in res._atoms:
                    out.add_atom(atom.name, atom.element, r, serial=atom.serial)
        for a1, a2 in self._bonds:
            out.add_bond(out.atom(a1.index), out.atom(a2.index))
        return out

    __copy__ = copy

    def __deepcopy__(self, memo):
        return self.copy()

    def subset(self, atom_indices):
        keep = set(int(i) for i in atom_indices)
        out = Topology()
        old_to_new = {}
        for chain in self._chains:
            new_chain = None
            for res in chain._residues:
                new_res = None
                for atom in res._atoms:
                    if atom.index in keep:
                        if new_chain is None:
                            new_chain = out.add_chain(chain.chain_id)
                        if new_res is None:
                            new_res = out.add_residue(res.name, new_chain, res.resSeq,
                                                      res.segment_id)
                        old_to_new[atom] = out.add_atom(atom.name, atom.element, new_res,
                                                        serial=atom.serial)
        for a1, a2 in self._bonds:
            if a1 in old_to_new and a2 in old_to_new:
                out.add_bond(old_to_new[a1], old_to_new[a2])
        return out

    def join(self, other, keep_resSeq=True):
        out = self.copy()
        for chain in other._chains:
            c = out.add_chain(chain.chain_id)
            for res in chain._residues:
                r = out.add_residue(res.name, c, res.resSeq if keep_resSeq else None,
                                    res.segment_id)
                for atom in res._atoms:
                    out.add_atom(atom.name, atom.element, r, serial=atom.serial)
        return out

    # -- data frame round trip -----------------------------------------
    def to_dataframe(self):
        import pandas as pd
        data = [(atom.serial, atom.name, atom.element.symbol, atom.residue.resSeq,
                 atom.residue.name, atom.residue.chain.index, atom.segment_id)
                for atom in self.atoms]
        atoms = pd.DataFrame(data, columns=["serial", "name", "element", "resSeq",
                                            "resName", "chainID", "segmentID"])
        bonds = np.array([(a.index, b.index, 0.0, 0.0) for (a, b) in self._bonds])
        return atoms, bonds

    @classmethod
    def from_dataframe(cls, atoms, bonds=None):
        out = cls()
        if "segmentID" not in atoms.columns:
            atoms = atoms.copy()
            atoms["segmentID"] = ""
        for ci in atoms["chainID"].unique():
            c = out.add_chain()
            chain_atoms = atoms[atoms["chainID"] == ci]
            prev_key = None
            r = None
            for _, row in chain_atoms.iterrows():
                key = (row["resSeq"], row["resName"], row["segmentID"])
                if key != prev_key:
                    seg = row["segmentID"]
                    r = out.add_residue(row["resName"], c, row["resSeq"],
                                        "" if seg is None else seg)
                    prev_key = key
                sym = row["element"]
                element = elem.get_by_symbol(sym) if sym else elem.virtual
                serial = row["serial"]
                serial = None if serial is None or (isinstance(serial, float) and np.isnan(serial)) \
                    else int(serial)
                out.add_atom(row["name"], element, r, serial=serial)
        if bonds is not None:
            for b in np.asarray(bonds).reshape(-1, np.asarray(bonds).shape[-1] if np.asarray(bonds).ndim > 1 else 1):
                if len(b) >= 2:
                    out.add_bond(out.atom(int(b[0])), out.atom(int(b[1])))
        return out

    # -- comparison ---------------------------------------------------
    def _signature(self):
        return (
            tuple((c.index,) for c in self._chains),
            tuple((r.index, r.name, r.resSeq, r.chain.index, r.segment_id)
                  for r in self.residues),
            tuple((a.index, a.name, a.element.symbol, a.residue.index) for a in self.atoms),
            tuple(sorted((min(a.index, b.index), max(a.index, b.index)) for a, b in self._bonds)),
        )

    def __eq__(self, other):
        if not isinstance(other, Topology):
            return False
        if self is other:
            return True
        return self._signature() == other._signature()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._signature())

    def __repr__(self):
        return "<oracle.mdtraj.Topology with %d chains, %d residues, %d atoms>" % (
            self.n_chains, self.n_residues, self.n_atoms)

    # -- selection ----------------------------------------------------
    def select(self, expression):
        pred = _compile_selection(expression)
        return np.array([a.index for a in self.atoms if pred(a)], dtype=int)


# ---------------------------------------------------------------------
# a small subset of the MDTraj atom-selection DSL
# ---------------------------------------------------------------------
_TOKEN = re.compile(r"\s*(\(|\)|==|!=|<=|>=|<|>|'[^']*'|\"[^\"]*\"|[^\s()=!<>]+)")

_FIELDS = {
    "symbol": lambda a: a.element.symbol, "element": lambda a: a.element.symbol,
    "type": lambda a: a.element.symbol,
    "name": lambda a: a.name,
    "resname": lambda a: a.residue.name, "resn": lambda a: a.residue.name,
    "resseq": lambda a: a.residue.resSeq, "residue": lambda a: a.residue.resSeq,
    "resi": lambda a: a.residue.resSeq,
    "resid": lambda a: a.residue.index,
    "index": lambda a: a.index,
    "chainid": lambda a: a.residue.chain.index,
    "mass": lambda a: a.element.mass,
}
_FLAGS = {
    "water": lambda a: a.residue.is_water, "waters": lambda a: a.residue.is_water,
    "is_water": lambda a: a.residue.is_water,
    "protein": lambda a: a.residue.is_protein, "is_protein": lambda a: a.residue.is_protein,
    "all": lambda a: True, "everything": lambda a: True,
    "none": lambda a: False, "nothing": lambda a: False,
}


def _literal(tok):
    if tok[0] in "'\"":
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        try:
            return float(tok)
        except ValueError:
            return tok


def _compile_selection(expression):
    tokens = _TOKEN.findall(expression)
    pos = [0]

    def peek():
        return tokens[pos[0]] if pos[0] < len(tokens) else None

    def take():
        tok = tokens[pos[0]]
        pos[0] += 1
        return tok

    def parse_or():
        left = parse_and()
        while peek() is not None and peek().lower() in ("or", "||"):
            take()
            right = parse_and()
            left = (lambda l, r: lambda a: l(a) or r(a))(left, right)
        return left

    def parse_and():
        left = parse_not()
        while peek() is not None and peek().lower() in ("and", "&&"):
            take()
            right = parse_not()
            left = (lambda l, r: lambda a: l(a) and r(a))(left, right)
        return left

    def parse_not():
        if peek() is not None and peek().lower() in ("not", "!"):
            take()
            inner = parse_not()
            return lambda a: not inner(a)
        return parse_atom()

    def parse_atom():
        tok = take()
        if tok == "(":
            inner = parse_or()
            if take() != ")":
                raise ValueError("unbalanced parentheses in selection")
            return inner
        low = tok.lower()
        if low in _FLAGS:
            return _FLAGS[low]
        if low in _FIELDS:
            getter = _FIELDS[low]
            nxt = peek()
            if nxt in ("==", "!=", "<", ">", "<=", ">="):
                op = take()
                val = _literal(take())
                ops = {"==": lambda x, y: x == y, "!=": lambda x, y: x != y,
                       "<": lambda x, y: x < y, ">": lambda x, y: x > y,
                       "<=": lambda x, y: x <= y, ">=": lambda x, y: x >= y}[op]
                return lambda a: ops(getter(a), val)
            first = _literal(take())
            if peek() is not None and peek().lower() == "to":
                take()
                last = _literal(take())
                return lambda a: first <= getter(a) <= last
            values = [first]
            while peek() is not None and peek().lower() not in ("and", "or", "not", ")", "&&", "||"):
                values.append(_literal(take()))
            return lambda a: getter(a) in values
        raise ValueError("unsupported selection token %r in the oracle shim" % tok)

    pred = parse_or()
    if pos[0] != len(tokens):
        raise ValueError("could not parse selection %r" % expression)
    return pred
