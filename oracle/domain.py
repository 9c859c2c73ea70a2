"""Oracle cell-domain analysis (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates ``RepairApi.computeDomainInErrorCells`` (``RepairApi.scala:479-675``) and the weak-label
pruning of ``ErrorModel._extract_error_cells_from`` (``errors.py:507-530``).
"""
import math

import numpy as np

from .table import cast_to_string


def _sort_key(v):
    # canonical candidate order = dictionary order of the discretised column
    return (0, v) if isinstance(v, (int, float)) else (1, str(v))


def compute_domain_in_error_cells(disc, row_id, err_cells, continuous, targets, fs, pairwise_stats,
                                  domain_stats, max_attrs_to_compute_domains, alpha, beta):
    """err_cells: list of (row_position, attribute, current_value_string).

    -> list of (row_position, attribute, current_value, domain) with domain = [(n, prob)] sorted
    by prob descending (ties: dictionary order of n), only entries with prob > beta.
    Targets that are continuous or have no correlated attribute get an empty domain.
    """
    assert len(disc.names) > 1
    assert 0 < max_attrs_to_compute_domains
    assert 0.0 <= alpha < 1.0 and 0.0 <= beta < 1.0 and alpha < beta
    assert targets
    row_count = disc.n_rows
    corr_attr_map = {k: v[:max_attrs_to_compute_domains] for k, v in pairwise_stats.items()}
    cells = [c for c in err_cells if c[1] in targets]  # :530-531
    disc_attrs = [a for a in disc.names if a != row_id]
    out = []
    for attribute, corr_with_scores in corr_attr_map.items():
        my_cells = [c for c in cells if c[1] == attribute]
        if not my_cells:
            continue
        use = attribute not in continuous and len(corr_with_scores) > 0  # :565
        # single-attribute bins of `attribute`: MAX(cnt) per value (:634-638)
        single = {}
        if attribute in fs.pos:
            ia = fs.pos[attribute]
            for vals, flags, cnt in fs.rows:
                if flags[ia] == 0 and all(f == 1 for j, f in enumerate(flags) if j != ia):
                    v = vals[ia]
                    if v is not None:
                        single[v] = max(single.get(v, cnt), cnt)
        # per correlated attribute: value of attr -> [(n, max(cnt - 1.0, 0.1))]   (:587-598)
        per_attr = []
        if use:
            for attr, _ in corr_with_scores:
                product_space = domain_stats[attr] * domain_stats[attribute]
                tau = int(alpha * (row_count // product_space))  # :572-576, integer division
                ia, it = fs.pos[attr], fs.pos[attribute]
                d = {}
                for vals, flags, cnt in fs.rows:
                    if vals[it] is not None and vals[ia] is not None and cnt > tau:
                        entry = (vals[it], max(float(cnt) - 1.0, 0.1))
                        lst = d.setdefault(vals[ia], [])
                        if entry not in lst:  # collect_set
                            lst.append(entry)
                per_attr.append((attr, d))
        for (r, _, cur) in my_cells:
            domain = None
            for attr, d in per_attr:
                v = disc.value(attr, r)
                rd = d.get(v) if v is not None else None
                # IF(ISNOTNULL(l.domain), CONCAT(l.domain, r.d), r.d)  -- CONCAT(x, NULL) = NULL
                if domain is not None:
                    domain = (domain + rd) if rd is not None else None
                else:
                    domain = rd
            scores = {}
            if domain:
                for n, c in domain:  # explode + join with single-attribute counts (:619-645)
                    cnt_n = single.get(n)
                    if cnt_n is None:
                        continue
                    s = math.exp(math.log(cnt_n / row_count) + math.log(c / cnt_n))
                    scores[n] = scores.get(n, 0.0) + s
            keys = sorted(scores.keys(), key=_sort_key)
            denom = 0.0
            for n in keys:
                denom += scores[n]
            dom = []
            for n in keys:
                prob = scores[n] / denom
                if prob > beta:
                    dom.append((n, prob))
            dom.sort(key=lambda t: -t[1])  # stable: ties keep dictionary order
            out.append((r, attribute, cur, dom))
    return out


def weak_labeled_cells(disc, cell_domains):
    """errors.py:517-524: a noisy cell is NOT an error when ``current_value = domain[0].n``."""
    weak = set()
    for (r, attribute, cur, dom) in cell_domains:
        if cur is None or not dom:
            continue
        n = dom[0][0]
        n_str = cast_to_string("str" if isinstance(n, str) else ("int" if isinstance(n, int) else "float"), n)
        if str(cur) == str(n_str):
            weak.add((r, attribute))
    return weak
