"""Small reader for the CPLEX-LP subset the reference accepts.

Mirrors the grammar of the reference's PEGTL parser (src/ILP/ILP_parser.cpp:25-160):
optional ``\\`` comment lines, ``Minimize``, signed objective terms ``[+-] [coef] [*] name``
(an optional trailing constant), ``Subject To``, one linear constraint per line with an
optional ``identifier:`` prefix, ``<=`` / ``>=`` / ``=``, integer right-hand side, then an
optional ``Bounds`` / ``Binaries`` / ``Generals`` tail that is ignored (all variables are
binary), and ``End``.  Variable indices are assigned in order of first appearance, the
objective first (src/ILP/ILP_input.cpp: add_new_variable is called from the objective
actions before any constraint is read).

Host-side input plumbing only: no part of the sweep runs here.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

LE, GE, EQ = 0, 1, 2  # inequality codes shared with bdd_b200.instances and the C-ABI tests

_NAME = r"[A-Za-z][A-Za-z0-9_\-/(){},#;\[\].']*"
_TERM = re.compile(r"\s*([+-])?\s*(\d+(?:\.\d*)?(?:[eE][+-]?\d+)?)?\s*\*?\s*(" + _NAME + r")")
_CONST = re.compile(r"\s*([+-])\s*(\d+(?:\.\d*)?(?:[eE][+-]?\d+)?)\s*$")
_SECTION = re.compile(r"^\s*(End|Bounds|Binaries|Generals|Coalesce)\b", re.IGNORECASE)


@dataclass
class Constraint:
    identifier: str
    variables: List[int]
    coefficients: List[int]
    ineq: int
    rhs: int


@dataclass
class ILP:
    objective: List[float] = field(default_factory=list)
    constant: float = 0.0
    var_names: List[str] = field(default_factory=list)
    var_index: Dict[str, int] = field(default_factory=dict)
    constraints: List[Constraint] = field(default_factory=list)

    def nr_variables(self) -> int:
        return len(self.var_names)

    def get_or_add_var(self, name: str) -> int:
        idx = self.var_index.get(name)
        if idx is None:
            idx = len(self.var_names)
            self.var_index[name] = idx
            self.var_names.append(name)
            self.objective.append(0.0)
        return idx


def _parse_terms(text: str) -> Tuple[List[Tuple[float, str]], float]:
    """Split ``+ 2 x - y + 3`` into [(2,'x'), (-1,'y')] and the trailing constant 3."""
    terms: List[Tuple[float, str]] = []
    pos = 0
    const = 0.0
    text = text.rstrip()
    while pos < len(text):
        m = _TERM.match(text, pos)
        if m is None or m.end() == pos:
            mc = _CONST.match(text[pos:])
            if mc is None:
                if text[pos:].strip() == "":
                    break
                raise ValueError(f"cannot parse LP expression near: {text[pos:pos+40]!r}")
            const = float(mc.group(2)) * (-1.0 if mc.group(1) == "-" else 1.0)
            break
        sign = -1.0 if m.group(1) == "-" else 1.0
        coef = float(m.group(2)) if m.group(2) is not None else 1.0
        terms.append((sign * coef, m.group(3)))
        pos = m.end()
    return terms, const


def parse_lp(text: str) -> ILP:
    """Parse an LP string (the reference's ``ILP_parser::parse_string``)."""
    lines = [ln for ln in text.replace("\r", "").split("\n")]
    lines = [ln for ln in lines if not ln.lstrip().startswith("\\")]
    ilp = ILP()
    i = 0
    while i < len(lines) and lines[i].strip() == "":
        i += 1
    if i >= len(lines) or lines[i].strip().lower() not in ("minimize", "minimise", "min"):
        raise ValueError("LP input must start with 'Minimize'")
    i += 1
    obj_text = []
    while i < len(lines) and lines[i].strip().lower() not in ("subject to", "st", "s.t.", "such that"):
        obj_text.append(lines[i])
        i += 1
    if i >= len(lines):
        raise ValueError("missing 'Subject To'")
    i += 1
    terms, const = _parse_terms(" ".join(obj_text))
    ilp.constant = const
    for coef, name in terms:
        ilp.objective[ilp.get_or_add_var(name)] += coef

    # constraints: a constraint may span several lines until its relation + rhs is seen
    pending = ""
    while i < len(lines):
        ln = lines[i]
        i += 1
        if ln.strip() == "":
            continue
        if pending == "" and _SECTION.match(ln):
            break
        pending += " " + ln
        m = re.search(r"(<=|>=|=<|=>|=)\s*([+-]?\s*\d+(?:\.\d*)?)\s*$", pending)
        if m is None:
            continue
        lhs = pending[: m.start()]
        rel = m.group(1)
        rhs_val = float(m.group(2).replace(" ", ""))
        ident = ""
        mi = re.match(r"\s*([^\s:]+)\s*:", lhs)
        if mi is not None:
            ident = mi.group(1)
            lhs = lhs[mi.end():]
        terms, const = _parse_terms(lhs)
        if rhs_val != int(rhs_val) or any(c != int(c) for c, _ in terms) or const != int(const):
            raise ValueError("constraints must have integer coefficients")
        merged: Dict[int, int] = {}
        order: List[int] = []
        for coef, name in terms:
            v = ilp.get_or_add_var(name)
            if v not in merged:
                merged[v] = 0
                order.append(v)
            merged[v] += int(coef)
        ineq = LE if rel in ("<=", "=<") else (GE if rel in (">=", "=>") else EQ)
        ilp.constraints.append(
            Constraint(ident, order, [merged[v] for v in order], ineq, int(rhs_val) - int(const))
        )
        pending = ""
    return ilp


def write_lp(ilp: ILP) -> str:
    """Inverse of :func:`parse_lp` (for round-trip tests and for feeding the reference)."""
    out = ["Minimize"]
    obj = " ".join(f"{'+' if c >= 0 else '-'} {abs(c)!r} {n}" for c, n in zip(ilp.objective, ilp.var_names))
    if ilp.constant != 0.0:
        obj += f" {'+' if ilp.constant >= 0 else '-'} {abs(ilp.constant)!r}"
    out.append(obj)
    out.append("Subject To")
    rel = {LE: "<=", GE: ">=", EQ: "="}
    for k, c in enumerate(ilp.constraints):
        lhs = " ".join(
            f"{'+' if a >= 0 else '-'} {abs(a)} {ilp.var_names[v]}" for a, v in zip(c.coefficients, c.variables)
        )
        out.append(f"{c.identifier or 'c' + str(k)}: {lhs} {rel[c.ineq]} {c.rhs}")
    out.append("End")
    return "\n".join(out) + "\n"
