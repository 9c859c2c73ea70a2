"""Host-side denial-constraint parser and classifier.

Syntax follows the reference (HoloClean style, ``DenialConstraints.scala:82-225``):
``t1&t2&EQ(t1.a,t2.a)&IQ(t1.b,t2.b)``, ``t1&EQ(t1.a,"c")&...`` and the sugar ``X->Y``.
Each parsed constraint is classified into the shape the CUDA detectors evaluate:

  CONST     every predicate compares a t1 attribute with a constant      -> dr_dc_const
  FD        EQ(a,a)+ & exactly one IQ(b,b), same attribute on both sides  -> dr_dc_fd_build/flag
  EQ_ONLY   only EQ(a,a) predicates: every row matches itself             -> all rows
  INEQ      EQ(a,a)* & exactly one LT(b,b) or GT(b,b): "another row of the key group has a larger /
            smaller b"                                                     -> dr_dc_fd_build + dr_dc_lt_flag
  OTHER     anything else (several inequalities, several IQs, cross-attribute predicates): not
            evaluated on the GPU in this version; raises NotImplementedError at detect time.
"""
import logging
import re
from collections import namedtuple

_logger = logging.getLogger("repair")

SIGNS = ("EQ", "IQ", "LT", "GT")
Predicate = namedtuple("Predicate", "sign left right right_is_attr")

_IDENT = re.compile(r"[a-zA-Z]+[a-zA-Z0-9]*\Z")


def _refs(pred):
    out = [pred.left]
    if pred.right_is_attr and pred.right != pred.left:
        out.append(pred.right)
    return out


def _split(text, sep):
    """Java ``String.split``: trailing empty pieces are dropped, but a string without the
    separator comes back whole (even when empty)."""
    pieces = text.split(sep)
    if len(pieces) > 1:
        while pieces and pieces[-1] == "":
            pieces.pop()
    return pieces


def parse_denial_constraint(text):
    tokens = [t.strip() for t in _split(text, "&")]
    two = len(tokens) >= 2 and _IDENT.match(tokens[0]) and _IDENT.match(tokens[1])
    one = not two and len(tokens) >= 1 and _IDENT.match(tokens[0])
    if not (two or one):
        if tokens:
            raise ValueError("Failed to parse an input string: '{}'".format(text))
        return []
    body = tokens[2:] if two else tokens[1:]
    if len(body) < 2:
        raise ValueError("At least two predicate candidates should be given, but {} candidates found: {}".format(
            len(body), text))
    t1 = re.escape(tokens[0])
    if two:
        rx = re.compile(r"(%s)\s*\(\s*%s\.(.*)\s*,\s*%s\.(.*)\s*\)\Z" % ("|".join(SIGNS), t1, re.escape(tokens[1])))
    else:
        rx = re.compile(r"(%s)\s*\(\s*%s\.(.*)\s*,\s*(.*)\)\Z" % ("|".join(SIGNS), t1))
    preds, illegal = [], []
    for piece in body:
        m = rx.match(piece)
        if m is None:
            illegal.append(piece)
        else:
            preds.append(Predicate(m.group(1), m.group(2).strip(), m.group(3).strip(), bool(two)))
    if illegal:
        raise ValueError("Illegal predicates found: " + ", ".join(illegal))
    return preds


def parse_fd_sugar(text):
    sides = [s.strip() for s in text.split("->")]
    sides = [s for s in sides if s]
    if len(sides) == 2:
        return [Predicate("EQ", sides[0], sides[0], True), Predicate("IQ", sides[1], sides[1], True)]
    if sides:
        raise ValueError("Failed to parse an input string: '{}'".format(text))
    return []


def load_statements(constraint_path, constraints):
    stmts = []
    if constraint_path is not None and constraint_path.strip():
        path = constraint_path[7:] if constraint_path.startswith("file://") else constraint_path
        try:
            with open(path) as f:
                stmts.extend(line.rstrip("\r\n") for line in f)
        except OSError:
            _logger.warning("Failed to load constrains from '{}'".format(constraint_path))
    if constraints is not None:
        stmts.extend(s.strip() for s in constraints.split(";") if s.strip())
    return stmts


def parse_and_verify(statements, table_attrs, input_name=""):
    """-> list of predicate lists whose attributes all exist in the table."""
    parsed = []
    for stmt in statements:
        try:
            try:
                preds = parse_denial_constraint(stmt)
            except Exception:
                preds = parse_fd_sugar(stmt)
        except Exception:
            _logger.warning("Illegal constraint format found: {}".format(stmt))
            continue
        parsed.append(preds)
    parsed = [p for p in parsed if p]
    known = set(table_attrs)
    missing = sorted({r for preds in parsed for p in preds for r in _refs(p)} - known)
    if missing:
        _logger.warning("Non-existent constraint attributes found in '{}': {}".format(input_name, ", ".join(missing)))
        parsed = [preds for preds in parsed if all(r in known for p in preds for r in _refs(p))]
    return parsed


def references(preds):
    out = []
    for p in preds:
        for r in _refs(p):
            if r not in out:
                out.append(r)
    return out


def classify(preds):
    """-> (shape, payload): see the module docstring."""
    if all(not p.right_is_attr for p in preds):
        return "CONST", list(preds)
    if all(p.right_is_attr and p.left == p.right for p in preds):
        eqs = [p.left for p in preds if p.sign == "EQ"]
        iqs = [p.left for p in preds if p.sign == "IQ"]
        if len(eqs) + len(iqs) == len(preds):
            if not iqs:
                return "EQ_ONLY", eqs
            if len(iqs) == 1 and eqs:
                return "FD", (list(dict.fromkeys(eqs)), iqs[0])
            if len(iqs) == 1 and not eqs:
                return "FD", ([], iqs[0])
        ineq = [p for p in preds if p.sign in ("LT", "GT")]
        if len(ineq) == 1 and len(eqs) + 1 == len(preds):
            return "INEQ", (list(dict.fromkeys(eqs)), ineq[0].sign, ineq[0].left)
    return "OTHER", list(preds)


def constant_literal(text):
    """The constant of a single-tuple predicate as SQL would read it: quoted -> string, else the
    raw token (numeric columns parse it as a number)."""
    t = text.strip()
    if len(t) >= 2 and t[0] == t[-1] and t[0] in "\"'":
        return t[1:-1], True
    return t, False
