"""Minimal chemical-element table for the oracle's mdtraj stand-in."""


class Element(object):
    def __init__(self, number, name, symbol, mass):
        self.number = number
        self.name = name
        self.symbol = symbol
        self.mass = mass

    def __repr__(self):
        return "<Element %s>" % self.symbol

    def __eq__(self, other):
        return isinstance(other, Element) and other.symbol == self.symbol

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.symbol)

    @property
    def atomic_number(self):
        return self.number


_TABLE = [
    (0, "virtual", "VS", 0.0), (1, "hydrogen", "H", 1.007947), (2, "helium", "He", 4.003),
    (3, "lithium", "Li", 6.9412), (5, "boron", "B", 10.811), (6, "carbon", "C", 12.01078),
    (7, "nitrogen", "N", 14.00672), (8, "oxygen", "O", 15.99943), (9, "fluorine", "F", 18.99840325),
    (11, "sodium", "Na", 22.98976928), (12, "magnesium", "Mg", 24.30506),
    (15, "phosphorus", "P", 30.9737622), (16, "sulfur", "S", 32.0655), (17, "chlorine", "Cl", 35.4532),
    (19, "potassium", "K", 39.09831), (20, "calcium", "Ca", 40.0784), (26, "iron", "Fe", 55.8452),
    (29, "copper", "Cu", 63.5463), (30, "zinc", "Zn", 65.4094), (34, "selenium", "Se", 78.963),
    (35, "bromine", "Br", 79.9041), (53, "iodine", "I", 126.904473),
]

_BY_SYMBOL = {}
for _n, _name, _sym, _mass in _TABLE:
    _BY_SYMBOL[_sym.upper()] = Element(_n, _name, _sym, _mass)

virtual = _BY_SYMBOL["VS"]
hydrogen = _BY_SYMBOL["H"]
carbon = _BY_SYMBOL["C"]
nitrogen = _BY_SYMBOL["N"]
oxygen = _BY_SYMBOL["O"]


def get_by_symbol(symbol):
    return _BY_SYMBOL[symbol.strip().upper()]
