"""Rule-based repairs (``setRepairByRules(True)``; reference ``model.py:583-673, 731-753, 928-953``
and ``DepGraph.scala:257-317``):

* nearest values -- an error cell whose current value is within ``model.rule.merge_threshold`` (update
  cost) of exactly one nearest value of the attribute's domain takes that value.  The costs are
  evaluated once per pair of dictionary entries on the host; the per-cell work (look-up, clearing the
  error bit, patching the column) runs on the device;
* functional dependencies -- a target ``y`` that a constraint ``X -> y`` declares dependent on ``x``
  is predicted by the map ``x -> y`` read off the clean cells (``dr_fd_map_build``), applied to the
  dirty-row tile by ``dr_tile_lut_fill``.

* regex structure -- an error cell of an attribute watched by a ``RegExErrorDetector`` is rebuilt from
  the detector's regex when its current value has the right shape (``regex_structure.py``); again one
  evaluation per dictionary entry, a repaired string the column has never held joins its dictionary.
"""
import logging
import re

import numpy as np

from . import constraints as DC
from .errors import ConstraintErrorDetector, RegExErrorDetector

_logger = logging.getLogger("repair")
_I32_MAX, _I32_MIN = 2 ** 31 - 1, -2 ** 31
_DICTIONARY = object()   # undo-list tag: a column dictionary that grew


def regex_repair_enabled(rm):
    return rm.repair_by_rules and not bool(rm._opt("model.rule.repair_by_regex.disabled"))


def _typed(col, code):
    """Dictionary entry as the Python value the reference's cost function sees."""
    v = col.dictionary[code]
    if col.kind == "str":
        return str(v)
    return int(v) if col.kind == "int" else float(v)


def _cost(cf, x, y):
    """UpdateCostFunction.compute (costs.py:33-34, 71-78): NULL unless both operands are truthy; a
    user function that raises yields NULL."""
    if not x or not y:
        return None
    try:
        c = cf._compute_impl(x, y)
    except Exception:
        return None
    return None if c is None else float(c)


def nearest_value_lut(col, cur_codes, domain_codes, cf, threshold):
    """-> int32[dict_size + 1] indexed by code + 1: the code an error cell with that current value is
    merged into, or -1 (model.py:598-616; NULL-cost candidates neither win nor count as runner-up)."""
    lut = np.full(col.dict_size + 1, -1, dtype=np.int32)
    strs = col.strings()
    cands = [(int(d), _typed(col, int(d))) for d in domain_codes]
    for c in cur_codes:
        cur = strs[int(c)]
        costs = sorted(((_cost(cf, cur, v), d) for d, v in cands if _cost(cf, cur, v) is not None),
                       key=lambda t: t[0])
        if len(costs) >= 2 and costs[0][0] <= threshold and costs[0][0] < costs[1][0]:
            lut[int(c) + 1] = costs[0][1]
    return lut


def _error_cell_values(engine, table, res, a):
    """-> (rows, current codes, int64 histogram of the current codes [dict_size + 1], all ranks)."""
    torch, ctx = engine.torch, engine.ctx
    col = table.by_name[a]
    rows = engine.bitmap_rows(res.bitmaps[a])
    n = int(rows.numel())
    cur = torch.empty(n, dtype=torch.int32, device=engine.device)
    ctx.gather(engine.dt.col(a), rows, n, cur)
    masked = torch.zeros(col.dict_size + 1, dtype=torch.int64, device=engine.device)
    ctx.scan_hist([cur], [col.dict_size], n, [None], masked)
    if engine.dist is not None:
        engine.dist.sum_(masked)
    return rows, cur, masked.cpu().numpy()


def _apply_lut(engine, table, res, a, rows, cur, lut, by_rules, undo, what):
    """Error cells whose current code maps to a code >= 0 in `lut` (indexed by code + 1) are decided:
    they leave the error bitmap and their new value is written into the resident table (the repair
    base; _repair_attrs, model.py:1326-1328)."""
    torch, ctx = engine.torch, engine.ctx
    col = table.by_name[a]
    n = int(rows.numel())
    lut_dev = torch.from_numpy(lut).to(engine.device)
    rep = torch.empty(n, dtype=torch.int32, device=engine.device)
    ctx.gather(lut_dev[1:], cur, n, rep)                        # cur = -1 reads lut[0]
    before = res.bitmaps[a].clone()
    ctx.bitmap_clear_rows(res.bitmaps[a], rows, (rep >= 0).to(torch.uint8), n)
    ctx.bitmap_andnot(before, res.bitmaps[a], engine.n_rows)    # the cells just decided
    p_rows = engine.bitmap_rows(before)
    m = int(p_rows.numel())
    old = torch.empty(m, dtype=torch.int32, device=engine.device)
    new = torch.empty(m, dtype=torch.int32, device=engine.device)
    ctx.gather(engine.dt.col(a), p_rows, m, old)
    ctx.gather(lut_dev[1:], old, m, new)
    ctx.scatter(engine.dt.col(a), p_rows, new, m)
    old_vals = None
    if col.continuous:
        old_vals = torch.empty(m, dtype=torch.float64, device=engine.device)
        ctx.gather(engine.dt.val(a), p_rows, m, old_vals, f64=True)
        new_vals = torch.from_numpy(np.asarray(col.dictionary, dtype=np.float64)[new.cpu().numpy()]).to(engine.device)
        ctx.scatter(engine.dt.val(a), p_rows, new_vals, m, f64=True)
    undo.append((a, p_rows, old, old_vals))
    by_rules.append((a, p_rows.cpu().numpy().astype(np.int64), old.cpu().numpy(), new.cpu().numpy()))
    res.n_cells[a] = res.n_cells.get(a, 0) - m
    engine._hist_cache.pop(a, None)                              # value counts changed
    _logger.info("[Repairing Phase] {} cells of '{}' repaired by {}".format(m, a, what))


def repair_by_regexs(rm, engine, table, res, by_rules, undo):
    """_repair_by_regexs (model.py:633-651; RepairApi.repairByRegularExpression :677-706)."""
    from .regex_structure import LexError, StructureRepair
    for det in rm.error_detectors:
        if not isinstance(det, RegExErrorDetector):
            continue
        a = det.attr
        if a not in res.target_columns or a not in res.bitmaps or (engine.dist is None and not res.n_cells.get(a)):
            continue
        col = table.by_name[a]
        if col.kind != "str":
            raise NotImplementedError("regex structure repair of a numeric attribute")
        try:
            fix = StructureRepair(det.regex)
        except (LexError, re.error) as e:
            _logger.warning("Repairing using regex '{}' (attr='{}') failed because: {}".format(det.regex, a, e))
            continue
        rows, cur, masked = _error_cell_values(engine, table, res, a)
        strs = col.strings()
        lut = np.full(col.dict_size + 1, -1, dtype=np.int32)
        fresh = []
        for c in np.nonzero(masked[1:] > 0)[0]:
            rep = fix(strs[int(c)])
            if rep is None:
                continue
            code = col.code_of(rep)
            if code < 0:                                        # a string the column has never held
                if rep not in fresh:
                    fresh.append(rep)
                code = col.dict_size + fresh.index(rep)
            lut[int(c) + 1] = code
        if not (lut >= 0).any():
            continue
        if fresh:
            undo.append((_DICTIONARY, col, col.dict_size))
            col.dictionary = np.concatenate([col.dictionary, np.array(fresh, dtype=object)])
            col._lut = None
        _apply_lut(engine, table, res, a, rows, cur, lut, by_rules, undo, "its regex")


def repair_by_nearest_values(rm, engine, table, res, by_rules, undo):
    """_repair_by_nearest_values (model.py:583-626)."""
    cf = rm.cf
    targets = [c for c in res.target_columns if c in cf.targets] if cf.targets else list(res.target_columns)
    threshold = float(rm._opt("model.rule.merge_threshold"))
    for a in targets:
        if a not in res.bitmaps or (engine.dist is None and not res.n_cells.get(a)):
            continue  # (sharded runs go on: the histograms below are collective)
        col = table.by_name[a]
        if col.dict_size == 0:
            continue
        rows, cur, masked = _error_cell_values(engine, table, res, a)
        # domain of the repair base = values that still occur once the error cells are blanked
        counts = np.asarray(engine.raw_value_counts(a), dtype=np.int64)
        counts = counts - masked[:len(counts)] if len(counts) >= len(masked) else \
            np.concatenate([counts, np.zeros(len(masked) - len(counts), dtype=np.int64)]) - masked
        lut = nearest_value_lut(col, np.nonzero(masked[1:] > 0)[0], np.nonzero(counts[1:] > 0)[0], cf, threshold)
        if (lut >= 0).any():
            _apply_lut(engine, table, res, a, rows, cur, lut, by_rules, undo, "nearest values")


def apply_rules(rm, engine, table, res):
    """_repair_by_rules (model.py:653-673): regex structure first, then nearest values.
    -> ([(attr, row positions, current codes, repaired codes)], undo list for `restore`)"""
    by_rules, undo = [], []
    if regex_repair_enabled(rm):
        repair_by_regexs(rm, engine, table, res, by_rules, undo)
    if rm._repair_by_nearest_values_enabled:
        repair_by_nearest_values(rm, engine, table, res, by_rules, undo)
    return by_rules, undo


def restore(engine, undo):
    """Puts the original values (and dictionaries) back."""
    for entry in reversed(undo):
        if entry[0] is _DICTIONARY:
            _, col, size = entry
            col.dictionary = col.dictionary[:size]
            col._lut = None
            continue
        a, rows, old, old_vals = entry
        engine.ctx.scatter(engine.dt.col(a), rows, old, int(rows.numel()))
        if old_vals is not None:
            engine.ctx.scatter(engine.dt.val(a), rows, old_vals, int(rows.numel()), f64=True)
        engine._hist_cache.pop(a, None)


def functional_deps(rm, table, target_columns):
    """_get_functional_deps + DepGraph.computeFunctionalDeps (model.py:735-753, DepGraph.scala:257-298):
    {y: sorted [x]} from the single ConstraintErrorDetector's statements, or None."""
    if not rm.repair_by_rules or bool(rm._opt("model.rule.repair_by_functional_deps.disabled")):
        return None
    ceds = [d for d in rm.error_detectors if isinstance(d, ConstraintErrorDetector)]
    if len(ceds) != 1:
        if ceds:
            _logger.warning("Multiple constraint classes not supported for detecting functional deps")
        return None
    ced = ceds[0]
    targets = [c for c in target_columns if c in ced.targets] if ced.targets else list(target_columns)
    stmts = DC.load_statements(ced.constraint_path, ced.constraints)
    fd = {}
    for preds in (DC.parse_and_verify(stmts, table.names, table.name) if stmts else []):
        if len(preds) != 2 or {p.sign for p in preds} != {"EQ", "IQ"}:
            continue
        if any(len(DC.references([p])) != 1 for p in preds):
            continue
        x = [p for p in preds if p.sign == "EQ"][0].left
        y = [p for p in preds if p.sign == "IQ"][0].left
        if y in targets and y not in fd.get(x, set()) and x not in fd.get(y, set()):
            fd.setdefault(y, set()).add(x)
    return {y: sorted(xs) for y, xs in fd.items()}


def build_fd_model(engine, table, res, x, y):
    """FunctionalDepModel (model.py:64-100) as a look-up table over x's dictionary.
    -> ("fd", x, device int32[dict_size(x) + 1] indexed by code + 1, {x code: y code})"""
    torch = engine.torch
    xc, yc = table.by_name[x], table.by_name[y]
    if yc.kind != "str":
        raise NotImplementedError("functional-dependency models only repair string attributes "
                                  "(the reference's FunctionalDepModel predicts strings)")
    lut = np.full(xc.dict_size + 1, -1, dtype=np.int32)
    mapping = {}
    if xc.kind == "str" and xc.dict_size:   # the reference's map is keyed by strings: other types never match
        lo = torch.full((xc.dict_size,), _I32_MAX, dtype=torch.int32, device=engine.device)
        hi = torch.full((xc.dict_size,), _I32_MIN, dtype=torch.int32, device=engine.device)
        targets = set(res.target_columns)
        engine.ctx.fd_map_build(engine.dt.col(x), res.bitmaps.get(x) if x in targets else None,
                                engine.dt.col(y), res.bitmaps.get(y) if y in targets else None,
                                engine.n_rows, xc.dict_size, lo, hi)
        if engine.dist is not None:
            engine.dist.min_(lo)
            engine.dist.max_(hi)
        lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
        single = (lo == hi) & (lo != _I32_MAX)
        lut[1:][single] = lo[single]
        mapping = {int(c): int(lo[c]) for c in np.nonzero(single)[0]}
    return ("fd", x, torch.from_numpy(lut).to(engine.device), mapping)


def resolve_prediction_order(models, target_columns):
    """_resolve_prediction_order (model.py:928-953)."""
    by_y = dict(models)
    ordered, waiting = [], list(target_columns)
    for y in target_columns:
        if by_y[y][0] != "fd":
            ordered.append((y, by_y[y]))
            waiting.remove(y)
    while waiting:
        before = len(waiting)
        for y in list(waiting):
            if by_y[y][1] not in waiting:
                ordered.append((y, by_y[y]))
                waiting.remove(y)
        assert len(waiting) < before
    return ordered
