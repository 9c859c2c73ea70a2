"""CPU oracle (TEST INFRASTRUCTURE ONLY): restatement of RegexStructureRepair
(src/main/scala/org/apache/spark/python/RegexStructureRepair.scala:28-139) and of the ANTLR grammar it
parses with (src/main/antlr4/org/apache/spark/python/RegexBase.g4).

The lexer is restated with ANTLR's rule: at every position the longest match wins, ties go to the
rule defined first (PATTERN, SYMBOL, RANGE, CONSTANT, the one-character operators, CHARACTER, NUMBER);
tab / CR / LF are skipped.  Patterns the reference's LEXER rejects raise ``ValueError`` (the reference
catches that and repairs nothing, RepairApi.scala:700-705).  Token sequences its PARSER would only
get through by error recovery (a SYMBOL / CHARACTER / NUMBER token, a dangling operator) raise
``NotImplementedError`` -- their outcome is not specified by the sources.
"""
import re

_CH = "[A-Za-z0-9]"
_PATTERN = r"\[(?:%s(?:-%s)?)+\]" % (_CH, _CH)
_SYMBOL = r"(?:%s|%s)" % (_PATTERN, _CH)
_RULES = [  # (name, regex) in the grammar's order
    ("PATTERN", _PATTERN),
    ("SYMBOL", _SYMBOL),
    ("RANGE", r"%s\{(?:[0-9]+|,[0-9]+|[0-9]+,|[0-9]+,[0-9]+)\}" % _SYMBOL),
    ("CONSTANT", r"[A-Za-z0-9 _\-%]+"),
    ("STAR", r"\*"), ("PLUS", r"\+"), ("MAYBE", r"\?"), ("ALTERNATION", r"\|"), ("ANY", r"\."),
    ("CARET", r"\^"), ("DOLLAR", r"\$"),
    ("CHARACTER", _CH), ("NUMBER", r"[0-9]+"), ("WS", r"[\t\r\n]+"),
]
_COMPILED = [(n, re.compile(r)) for n, r in _RULES]


def _longest(rx, text, pos):
    """Longest match of `rx` at `pos` (Python's alternation is ordered, ANTLR's is not: try every
    alternative by anchoring candidates of decreasing length)."""
    m = rx.match(text, pos)
    if not m:
        return 0
    best = m.end() - pos
    for end in range(len(text), pos + best, -1):
        if rx.fullmatch(text, pos, end):
            return end - pos
    return best


def lex(pattern):
    tokens, pos = [], 0
    while pos < len(pattern):
        best_name, best_len = None, 0
        for name, rx in _COMPILED:
            n = _longest(rx, pattern, pos)
            if n > best_len:
                best_name, best_len = name, n
        if best_len == 0:
            raise ValueError("token recognition error at: '{}'".format(pattern[pos]))
        if best_name != "WS":
            tokens.append((best_name, pattern[pos:pos + best_len]))
        pos += best_len
    return tokens


def parse(pattern):
    """RegexParser.parse (:73-83) -> [(kind, text)] with kind in Pattern / Constant / Other."""
    tokens = lex(pattern)
    out, i = [], 0
    if i < len(tokens) and tokens[i][0] == "CARET":
        out.append(("Other", tokens[i][1]))
        i += 1
    body_end = len(tokens)
    tail = []
    if body_end > i and tokens[body_end - 1][0] == "DOLLAR":
        tail = [("Other", tokens[body_end - 1][1])]
        body_end -= 1
    have_expr, after_alt = False, False
    for name, text in tokens[i:body_end]:
        if name in ("CONSTANT", "RANGE", "PATTERN", "ANY"):
            if name == "CONSTANT":
                out.append(("Constant", text))
            elif name == "RANGE":
                out.append(("Pattern", text))
            have_expr, after_alt = True, False
        elif name in ("STAR", "PLUS", "MAYBE"):
            if not have_expr or after_alt:
                raise NotImplementedError("operator without an operand in '{}'".format(pattern))
        elif name == "ALTERNATION":
            if not have_expr or after_alt:
                raise NotImplementedError("operator without an operand in '{}'".format(pattern))
            after_alt = True
        else:
            raise NotImplementedError("token {} '{}' is outside the grammar's expression rule".format(name, text))
    if not have_expr or after_alt:
        raise NotImplementedError("no complete expression in '{}'".format(pattern))
    return out + tail


class RegexStructureRepair:
    """RegexStructureRepair (:95-139): every ranged pattern becomes a capture group, every constant
    ``.{1,len}``; a value that matches is rebuilt from its captured groups and the constants."""

    def __init__(self, pattern):
        self.tokens = parse(pattern)
        parts = []
        for kind, text in self.tokens:
            parts.append("(" + text + ")" if kind == "Pattern" else
                         ".{1,%d}" % len(text) if kind == "Constant" else text)
        self.regex = re.compile("".join(parts))

    def __call__(self, s):
        if s is None:
            return None
        m = self.regex.search(s)
        if m is None:
            return None
        out, g = [], 0
        for kind, text in self.tokens:
            if kind == "Pattern":
                g += 1
                out.append(m.group(g))
            elif kind == "Constant":
                out.append(text)
        return "".join(out)
