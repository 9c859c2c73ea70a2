"""Regex structure repair (``model.rule.repair_by_regex``; reference
``src/main/scala/org/apache/spark/python/RegexStructureRepair.scala:95-139`` with the token grammar
``src/main/antlr4/org/apache/spark/python/RegexBase.g4``).

A detector regex such as ``^[0-9]{1,3} patients$`` is read as a sequence of ranged character classes
(kept as capture groups) and literal runs (relaxed to "any 1..len characters"); a dirty value that
matches the relaxed regex is rebuilt from its captured groups and the literal runs:
``32 patxxnts`` -> ``32 patients``.  The function is evaluated once per dictionary entry on the host.
"""
import re

_ALNUM = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
_LITERAL = _ALNUM + " _-%"
_OPERATORS = {"*": "STAR", "+": "PLUS", "?": "MAYBE", "|": "ALT", ".": "ANY", "^": "CARET", "$": "DOLLAR"}


class LexError(ValueError):
    """The pattern holds a character sequence no token of the grammar covers (the reference logs a
    warning and repairs nothing)."""


def _class_end(p, i):
    """End of a bracket class ``[a-z0-9]`` starting at i, or -1.  Members are single alphanumerics or
    alphanumeric ranges; the longest well-formed class wins."""
    if i >= len(p) or p[i] != "[":
        return -1
    j, n_members, best = i + 1, 0, -1
    # members: X or X-Y; greedy with one character of look-ahead is exact for this grammar because a
    # '-' can only continue a range
    while j < len(p) and p[j] in _ALNUM:
        if j + 2 < len(p) and p[j + 1] == "-" and p[j + 2] in _ALNUM:
            j += 3
        else:
            j += 1
        n_members += 1
        if j < len(p) and p[j] == "]":
            best = j + 1
    return best if n_members else -1


def _symbol_end(p, i):
    e = _class_end(p, i)
    if e > 0:
        return e
    return i + 1 if i < len(p) and p[i] in _ALNUM else -1


def _digits_end(p, i):
    j = i
    while j < len(p) and p[j].isdigit() and p[j] in _ALNUM:
        j += 1
    return j


def _range_end(p, i):
    """SYMBOL '{' (N | ,N | N, | N,M) '}'"""
    s = _symbol_end(p, i)
    if s < 0 or s >= len(p) or p[s] != "{":
        return -1
    j = s + 1
    a = _digits_end(p, j)
    if a > j:                                   # N...
        if a < len(p) and p[a] == "}":
            return a + 1
        if a < len(p) and p[a] == ",":
            b = _digits_end(p, a + 1)
            return b + 1 if b < len(p) and p[b] == "}" else -1
        return -1
    if j < len(p) and p[j] == ",":              # ,N
        b = _digits_end(p, j + 1)
        return b + 1 if b > j + 1 and b < len(p) and p[b] == "}" else -1
    return -1


def _literal_end(p, i):
    j = i
    while j < len(p) and p[j] in _LITERAL:
        j += 1
    return j


def tokenize(pattern):
    """Longest token at every position; ties go to the rule the grammar lists first (class, symbol,
    range, literal run, operators).  -> [(kind, text)]"""
    out, i = [], 0
    while i < len(pattern):
        if pattern[i] in "\t\r\n":
            i += 1
            continue
        cands = [("CLASS", _class_end(pattern, i)), ("SYMBOL", _symbol_end(pattern, i)),
                 ("RANGE", _range_end(pattern, i)), ("LITERAL", _literal_end(pattern, i))]
        if pattern[i] in _OPERATORS:
            cands.append((_OPERATORS[pattern[i]], i + 1))
        kind, end = None, i
        for k, e in cands:
            if e > end:
                kind, end = k, e
        if kind is None:
            raise LexError("token recognition error at: '{}'".format(pattern[i]))
        out.append((kind, pattern[i:end]))
        i = end
    return out


def structure(pattern):
    """-> [("group" | "literal" | "anchor", text)] or raises: LexError for a pattern the grammar's
    lexer rejects, NotImplementedError for one its parser would only get through by error recovery."""
    toks = tokenize(pattern)
    head = [("anchor", "^")] if toks and toks[0][0] == "CARET" else []
    body = toks[len(head):]
    tail = []
    if body and body[-1][0] == "DOLLAR":
        tail, body = [("anchor", "$")], body[:-1]
    parts, operand, pending_alt = [], False, False
    for kind, text in body:
        if kind in ("LITERAL", "RANGE", "CLASS", "ANY"):
            if kind == "LITERAL":
                parts.append(("literal", text))
            elif kind == "RANGE":
                parts.append(("group", text))
            operand, pending_alt = True, False
        elif kind in ("STAR", "PLUS", "MAYBE", "ALT") and operand and not pending_alt:
            pending_alt = kind == "ALT"
        else:
            raise NotImplementedError("regex '{}' is outside the structure-repair grammar at '{}'".format(pattern, text))
    if not operand or pending_alt:
        raise NotImplementedError("regex '{}' is outside the structure-repair grammar".format(pattern))
    return head + parts + tail


class StructureRepair:
    def __init__(self, pattern):
        self.parts = structure(pattern)
        relaxed = "".join("(" + t + ")" if k == "group" else ".{1,%d}" % len(t) if k == "literal" else t
                          for k, t in self.parts)
        self.relaxed = re.compile(relaxed)

    def __call__(self, value):
        if value is None:
            return None
        m = self.relaxed.search(value)
        if m is None:
            return None
        groups = iter(m.groups())
        return "".join(next(groups) if k == "group" else t for k, t in self.parts if k != "anchor")
