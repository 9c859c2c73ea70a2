"""Device pipeline: the B200 replacement for ``ErrorModel.detect`` (errors.py:545-582) and the
inference half of ``RepairModel._run`` (model.py:1288-1408).

Everything that touches table-sized data is a call into ``libb200repair.so``; PyTorch tensors are
only the buffers those calls read and write (plus ``torch.distributed`` for the one exchange step
of the row-sharded multi-GPU path).  Host code here handles dictionaries, the tiny count tensors
and model bookkeeping.
"""
import logging
import re

import time

import numpy as np

from . import constraints as DC
from . import stats_host as SH
from ._native import DR_OP, Context
from .table import DeviceTable

_logger = logging.getLogger("repair")

MAX_FD_KEY_SPACE = 1 << 27      # int32 lo + hi tables: 1 GiB
PRESENCE_SAMPLE_ROWS = 1 << 20  # rows used for distinct-pair lower bounds


class Dist:
    """The single exchange step of the sharded path: all-reduce of count tensors."""

    def __init__(self):
        import torch.distributed as td
        self.td = td
        self.rank = td.get_rank()
        self.world = td.get_world_size()

    def sum_(self, t):
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)

    def min_(self, t):
        self.td.all_reduce(t, op=self.td.ReduceOp.MIN)

    def max_(self, t):
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)


class DetectResult:
    def __init__(self):
        self.bitmaps = {}            # attr -> device int32 words (noisy / error cells)
        self.noisy_columns = []
        self.target_columns = []
        self.pairwise_stats = {}
        self.domain_stats = {}
        self.disc_attrs = []
        self.n_cells = {}
        self.n_cells_detected = {}
        self.weak_removed = 0


class Engine:
    def __init__(self, table, device_index=0, dist=None, device_table=None):
        import torch
        self.torch = torch
        self.ctx = Context(device_index)
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        self.table = table
        self.dt = device_table if device_table is not None else DeviceTable(table, self.device, ctx=self.ctx)
        self.dist = dist
        self.n_rows = table.n_rows
        self.n_words = (self.dt.n_pad + 31) // 32
        self._hist_cache = {}
        self.disc_cols = {}      # attr -> device int32 column of the discretised table
        self.disc_dom = {}       # attr -> domain size of that column
        self.timings = {}
        self.trace = None        # list of (label, seconds since the previous mark) when tracing

    # ------------------------------------------------------------------------------------------
    @property
    def n_rows_global(self):
        return self.table.n_rows_global

    def pinned_i32(self, n):
        """Page-locked int32 staging of at least n elements, kept for the life of the engine
        (cudaHostAlloc of a few hundred MB per run would cost more than the copy it serves)."""
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < n:
            buf = self.torch.empty(int(n * 1.25) + 1024, dtype=self.torch.int32, pin_memory=True)
            self._pinned = buf
        return buf[:n]

    def mark(self, label):
        """Tracing aid (bench.py --trace): device-synchronised wall-clock split points."""
        if self.trace is None:
            return
        self.torch.cuda.synchronize()
        now = time.perf_counter()
        self.trace.append((label, now - getattr(self, "_t_mark", now)))
        self._t_mark = now

    def new_bitmap(self):
        return self.torch.zeros(self.n_words, dtype=self.torch.int32, device=self.device)

    def _sync(self):
        self.torch.cuda.current_stream().synchronize()

    def bitmap_rows(self, bitmap, n=None, out=None, count=None):
        """Ascending row indices (device int32 tensor) of the set bits (written into `out` if given).
        With `count` (the popcount, e.g. from ctx.bitmap_count_many) there is no host round trip."""
        n = self.n_rows if n is None else n
        if count is not None:
            rows = out if out is not None else \
                self.torch.empty(max(count, 1), dtype=self.torch.int32, device=self.device)
            assert rows.numel() >= count
            if count:
                self.ctx.bitmap_to_rows_async(bitmap, n, rows, count)
            return rows[:count]
        cnt = self.ctx.bitmap_count(bitmap, n)
        rows = out if out is not None else \
            self.torch.empty(max(cnt, 1), dtype=self.torch.int32, device=self.device)
        assert rows.numel() >= cnt
        if cnt:
            self.ctx.bitmap_rows_after_count(bitmap, n, rows, cnt)  # reuses the block offsets of the count
        return rows[:cnt]

    # ---- a7: discretisation ------------------------------------------------------------------
    def discretize(self, discrete_thres):
        """convertToDiscretizedTable (RepairApi.scala:126-169) -> domain_stats; fills disc_cols."""
        assert 2 <= discrete_thres < 65536
        domain_stats = {}
        self.disc_cols, self.disc_dom = {}, {}
        for c in self.table.columns:
            ndv = c.dict_size
            domain_stats[c.name] = ndv
            if c.continuous:
                out = self.torch.full((self.dt.n_pad,), -1, dtype=self.torch.int32, device=self.device)
                if ndv > 0:
                    vmin, den = SH.discretize_params(c.kind, c.dictionary[0], c.dictionary[-1])
                    self.ctx.discretize(self.dt.val(c.name), self.n_rows, vmin, den, discrete_thres, out)
                self.disc_cols[c.name] = out
                self.disc_dom[c.name] = discrete_thres + 1
            elif 1 < ndv <= discrete_thres:
                self.disc_cols[c.name] = self.dt.col(c.name)
                self.disc_dom[c.name] = ndv
            else:
                _logger.warning("'{}' dropped because of its unsuitable domain (size={})".format(c.name, ndv))
        return domain_stats

    # ---- histograms ----------------------------------------------------------------------------
    def scan_hist(self, attrs, null_bitmaps):
        """Fused NULL scan + histogram over discretised columns `attrs`; caches hist per attr."""
        if not attrs:
            return
        for i in range(0, len(attrs), 64):
            part = attrs[i:i + 64]
            dom = [self.disc_dom[a] for a in part]
            hist = self.torch.zeros(sum(d + 1 for d in dom), dtype=self.torch.int64, device=self.device)
            self.ctx.scan_hist([self.disc_cols[a] for a in part], dom, self.n_rows,
                               [null_bitmaps.get(a) for a in part], hist)
            if self.dist is not None:
                self.dist.sum_(hist)
            h = hist.cpu().numpy()
            off = 0
            for a, d in zip(part, dom):
                self._hist_cache[a] = h[off:off + d + 1].copy()
                off += d + 1

    def raw_value_counts(self, attr):
        """int64[dict_size + 1] counts of the RAW column (slot 0 = NULL)."""
        col = self.table.by_name[attr]
        if attr in self._hist_cache and not col.continuous and attr in self.disc_cols:
            return self._hist_cache[attr]
        hist = self.torch.zeros(col.dict_size + 1, dtype=self.torch.int64, device=self.device)
        self.ctx.scan_hist([self.dt.col(attr)], [col.dict_size], self.n_rows, [None], hist)
        if self.dist is not None:
            self.dist.sum_(hist)
        return hist.cpu().numpy()

    # ---- detectors -----------------------------------------------------------------------------
    def _or_rows_into(self, row_bitmap, attrs, bitmaps):
        for a in attrs:
            if a not in bitmaps:
                bitmaps[a] = self.new_bitmap()
            self.ctx.bitmap_or(bitmaps[a], row_bitmap, self.n_rows)

    def detect_null(self, targets, bitmaps, fused):
        """NullErrorDetector: columns of the discretised table get their bits from the fused scan
        (`fused` collects them); the others take a NULL-only LUT scan of the raw codes."""
        for a in self.table.names:
            if a not in targets:
                continue
            if a not in bitmaps:
                bitmaps[a] = self.new_bitmap()
            col = self.table.by_name[a]
            if a in self.disc_cols and not col.continuous:
                fused[a] = bitmaps[a]
            else:
                self.ctx.lut_scan(self.dt.col(a), self.n_rows, None, 0, bitmaps[a])

    def detect_regex(self, attr, regex, targets, bitmaps):
        if attr not in targets or regex is None or regex.strip() == "" or attr not in self.table.by_name:
            return
        col = self.table.by_name[attr]
        pat = re.compile(regex)
        lut = np.fromiter((0 if pat.search(s) is not None else 1 for s in col.strings()), dtype=np.uint8,
                          count=col.dict_size)
        d_lut = self.torch.from_numpy(lut).to(self.device) if col.dict_size else None
        if attr not in bitmaps:
            bitmaps[attr] = self.new_bitmap()
        self.ctx.lut_scan(self.dt.col(attr), self.n_rows, d_lut, col.dict_size, bitmaps[attr])

    def domain_values_regex(self, attr, values, autofill, min_count_thres):
        """DomainValues._detect_impl (errors.py:112-129) -> regex or None (continuous attribute)."""
        if attr not in self.table.by_name:
            return "$^" if not values or autofill else "({})".format("|".join(values))
        col = self.table.by_name[attr]
        if col.continuous:
            return None
        domain_values = [] if autofill else list(values)
        if autofill:
            counts = self.raw_value_counts(attr)
            strs = col.strings()
            filled = [strs[c] for c in range(col.dict_size) if counts[c + 1] > min_count_thres]
            if filled:
                domain_values = filled
        return "({})".format("|".join(domain_values)) if domain_values else "$^"

    def detect_constraints(self, constraint_path, constraints, targets, bitmaps):
        stmts = DC.load_statements(constraint_path, constraints)
        if not stmts:
            return
        pred_lists = DC.parse_and_verify(stmts, [self.table.row_id] + self.table.names, self.table.name)
        for preds in pred_lists:
            attrs = [r for r in DC.references(preds) if r in targets]
            if not attrs:
                continue
            shape, payload = DC.classify(preds)
            rowmask = self.new_bitmap()
            if shape == "CONST":
                cols, ops, args = [], [], []
                for p in payload:
                    col = self.table.by_name[p.left]
                    lit, quoted = DC.constant_literal(p.right)
                    cols.append(self.dt.col(p.left))
                    ops.append(DR_OP[p.sign])
                    if p.sign in ("EQ", "IQ"):
                        args.append(col.code_of(lit))
                    else:
                        lo, hi = col.rank_bounds(lit)
                        args.append(lo if p.sign == "LT" else hi)
                self.ctx.dc_const(cols, ops, args, self.n_rows, rowmask)
            elif shape == "FD":
                keys, b = payload
                self._flag_by_key_group(keys, self.dt.col(b), 0, rowmask)
            elif shape == "INEQ":
                keys, sign, b = payload
                x = self.dt.col(b)
                if sign == "GT":   # "some row has a smaller b" = "some row has a larger reversed code"
                    dom = self.table.by_name[b].dict_size
                    x = self.torch.where(x >= 0, dom - 1 - x, x)
                self._flag_by_key_group(keys, x, 1, rowmask)
            elif shape == "EQ_ONLY":
                rowmask.fill_(-1)
            else:
                raise NotImplementedError(
                    "denial constraint shape not supported on the GPU path yet: {}".format(
                        " & ".join("{}({},{})".format(p.sign, p.left, p.right) for p in preds)))
            self._or_rows_into(rowmask, attrs, bitmaps)

    def _flag_by_key_group(self, keys, x, mode, rowmask):
        """Rows whose NULL-safe key group (attributes `keys`) holds two distinct x (mode 0: EQ.. & IQ(x))
        or a larger x than their own (mode 1: EQ.. & LT(x)).  Per-key min / max of x + 1: in direct tables
        over the mixed-radix key space when it is small enough, else in a hash table keyed by the key."""
        torch = self.torch
        strides, space = [], 1
        for k in keys:
            strides.append(space)
            space *= self.table.by_name[k].dict_size + 1
        key_cols = [self.dt.col(k) for k in keys] or [x]
        if not keys:
            strides = [0]
        if space <= MAX_FD_KEY_SPACE:
            lo = torch.full((space,), 2 ** 31 - 1, dtype=torch.int32, device=self.device)
            hi = torch.full((space,), -2 ** 31, dtype=torch.int32, device=self.device)
            self.ctx.dc_fd_build(key_cols, strides, x, self.n_rows, space, lo, hi)
            if self.dist is not None:
                self.dist.min_(lo)
                self.dist.max_(hi)
            if mode == 0:
                self.ctx.dc_fd_flag(key_cols, strides, self.n_rows, space, lo, hi, rowmask)
            else:
                self.ctx.dc_lt_flag(key_cols, strides, x, self.n_rows, space, hi, rowmask)
            return
        if space >= 2 ** 62:
            raise NotImplementedError("denial constraint key space {} does not fit a 64-bit key".format(space))
        if self.dist is not None:
            raise NotImplementedError("denial constraint key space {} needs the hash-table path, which is "
                                      "single-GPU (tables of different ranks cannot be all-reduced)".format(space))
        cap = 1 << max(1, (2 * max(self.n_rows, 1) - 1).bit_length())
        tkeys = torch.full((cap,), -1, dtype=torch.int64, device=self.device)
        lo = torch.full((cap,), 2 ** 31 - 1, dtype=torch.int32, device=self.device)
        hi = torch.full((cap,), -2 ** 31, dtype=torch.int32, device=self.device)
        self.ctx.dc_hash_build(key_cols, strides, x, self.n_rows, cap, tkeys, lo, hi)
        self.ctx.dc_hash_flag(key_cols, strides, x if mode == 1 else None, mode, self.n_rows, cap, tkeys, lo, hi,
                              rowmask)

    def detect_outliers(self, targets, bitmaps, approx_enabled=False):
        for a in self.table.continuous_attrs:
            if a not in targets:
                continue
            if self.dist is not None:
                raise NotImplementedError("exact quartiles are single-GPU in this version")
            q1, q3, n = self.ctx.quartiles(self.dt.val(a), self.n_rows)
            if n == 0:
                continue
            lower, upper = q1 - 1.5 * (q3 - q1), q3 + 1.5 * (q3 - q1)
            if a not in bitmaps:
                bitmaps[a] = self.new_bitmap()
            self.ctx.range_flag(self.dt.val(a), self.n_rows, lower, upper, bitmaps[a])

    def bitmaps_from_cells(self, positions, attrs):
        """User-supplied error cells (setErrorCells) -> bitmaps, built on the host (tiny)."""
        out = {}
        for a in sorted(set(attrs)):
            pos = np.asarray([p for p, x in zip(positions, attrs) if x == a], dtype=np.int64)
            words = np.zeros(self.n_words, dtype=np.uint32)
            np.bitwise_or.at(words, pos >> 5, (np.uint32(1) << (pos & 31).astype(np.uint32)))
            out[a] = self.torch.from_numpy(words.view(np.int32)).to(self.device)
        return out

    # ---- a8: pair statistics -----------------------------------------------------------------
    def _pair_layout(self, pairs, bits):
        offs = [0]
        for x, y in pairs:
            n = (self.disc_dom[x] + 1) * (self.disc_dom[y] + 1)
            offs.append(offs[-1] + ((n + 31) // 32 if bits else n))
        return offs

    def pair_nnz_lower_bounds(self, pairs):
        """distinct-pair counts on a row sample (exact when the sample covers the table)."""
        attrs = list(dict.fromkeys(a for p in pairs for a in p))
        idx = {a: i for i, a in enumerate(attrs)}
        offs = self._pair_layout(pairs, bits=True)
        bits = self.torch.zeros(max(offs[-1], 1), dtype=self.torch.int32, device=self.device)
        block_rows = 256
        n_blocks = min((self.n_rows + block_rows - 1) // block_rows, PRESENCE_SAMPLE_ROWS // block_rows)
        exact = n_blocks * block_rows >= self.n_rows and self.dist is None
        if len(attrs) > 64:
            raise NotImplementedError("more than 64 discretised attributes")
        self.ctx.pair_presence([self.disc_cols[a] for a in attrs], [self.disc_dom[a] for a in attrs],
                               [idx[x] for x, _ in pairs], [idx[y] for _, y in pairs], offs, self.n_rows,
                               block_rows, n_blocks, bits)
        words = bits.cpu().numpy().view(np.uint32)
        if self.dist is not None:
            un = self.torch.from_numpy(np.unpackbits(words.view(np.uint8), bitorder="little")).to(self.device)
            self.dist.max_(un)
            words = np.packbits(un.cpu().numpy(), bitorder="little").view(np.uint32)
        nnz = {}
        for q, (x, y) in enumerate(pairs):
            w = words[offs[q]:offs[q + 1]]
            nnz[frozenset((x, y))] = int(np.unpackbits(w.view(np.uint8)).sum())
        return nnz, exact

    def pair_tables(self, pairs):
        """Exact co-occurrence tables {(x, y): int64[dom_x+1, dom_y+1]} for `pairs`."""
        if not pairs:
            return {}
        attrs = list(dict.fromkeys(a for p in pairs for a in p))
        if len(attrs) > 64:
            raise NotImplementedError("more than 64 discretised attributes")
        idx = {a: i for i, a in enumerate(attrs)}
        offs = self._pair_layout(pairs, bits=False)
        out = self.torch.zeros(offs[-1], dtype=self.torch.int64, device=self.device)
        self.ctx.cooc([self.disc_cols[a] for a in attrs], [self.disc_dom[a] for a in attrs],
                      [idx[x] for x, _ in pairs], [idx[y] for _, y in pairs], offs, self.n_rows, out)
        if self.dist is not None:
            self.dist.sum_(out)
        h = out.cpu().numpy()
        return {(x, y): h[offs[q]:offs[q + 1]].reshape(self.disc_dom[x] + 1, self.disc_dom[y] + 1).copy()
                for q, (x, y) in enumerate(pairs)}

    def compute_attr_stats(self, targets, domain_stats, attr_freq_thr, pairwise_thr, max_attrs):
        """computeAttrStats (RepairApi.scala:396-477) -> (pairwise_stats, tables, having)."""
        disc_attrs = list(self.disc_cols.keys())
        cands = SH.candidate_pairs(targets, disc_attrs)
        scoring = [t for t in targets if len(cands[t]) > max_attrs]

        def uniq(pairs):
            seen, out = set(), []
            for p in pairs:
                k = frozenset(p)
                if k not in seen:
                    seen.add(k)
                    out.append(p)
            return out

        selected = {t: list(cands[t]) for t in targets if t not in scoring}
        tables = {}
        if scoring:
            all_scored = uniq([p for t in scoring for p in cands[t]])
            lower, exact = self.pair_nnz_lower_bounds(all_scored)
            if exact:
                for t in scoring:
                    selected[t] = SH.select_scored(cands[t], lower, domain_stats, pairwise_thr, max_attrs)
            else:
                und = {t: SH.undecided(cands[t], lower, domain_stats, pairwise_thr) for t in scoring}
                need = uniq([p for t in scoring for p in und[t]] + [p for t in selected for p in selected[t]])
                tables = self.pair_tables(need)
                nnz = {frozenset(k): int(np.count_nonzero(v)) for k, v in tables.items()}
                for t in scoring:
                    selected[t] = SH.select_scored(und[t], nnz, domain_stats, pairwise_thr, max_attrs)
        pairs = [p for t in targets for p in selected[t]]
        missing = [p for p in uniq(pairs) if p not in tables and (p[1], p[0]) not in tables]
        tables.update(self.pair_tables(missing))
        having = SH.having_threshold(self.n_rows_global, attr_freq_thr)
        need_hist = [a for a in disc_attrs if a not in self._hist_cache]
        self.scan_hist(need_hist, {})
        stats = SH.pairwise_entropies(self.n_rows_global, self._hist_cache, tables, pairs, domain_stats, having)
        for t in targets:
            stats.setdefault(t, [])
        return stats, tables, having

    # ---- a9: weak-label domain analysis --------------------------------------------------------
    def prune_weak_labels(self, res, tables, having, continuous, max_attrs_domains, alpha, beta):
        removed = 0
        for t in res.target_columns:
            corr = res.pairwise_stats.get(t, [])[:max_attrs_domains]
            if t in continuous or not corr or t not in res.bitmaps:
                continue
            rows = self.bitmap_rows(res.bitmaps[t])
            n = int(rows.numel())
            if n == 0:
                continue
            dom_t = self.disc_dom[t]
            hist_t = SH.apply_having(self._hist_cache[t], having).astype(np.int64)
            cooc, dom_c, taus, ccols = [], [], [], []
            for a, _ in corr:
                tab = tables[(t, a)].T if (t, a) in tables else tables[(a, t)]
                tab = SH.apply_having(np.ascontiguousarray(tab), having).astype(np.int64)  # [dom_a+1, dom_t+1]
                cooc.append(self.torch.from_numpy(np.ascontiguousarray(tab)).to(self.device))
                dom_c.append(self.disc_dom[a])
                taus.append(SH.tau_for(alpha, self.n_rows_global, res.domain_stats[a], res.domain_stats[t]))
                ccols.append(self.disc_cols[a])
            top1 = self.torch.empty(n, dtype=self.torch.int32, device=self.device)
            prob = self.torch.empty(n, dtype=self.torch.float64, device=self.device)
            weak = self.torch.empty(n, dtype=self.torch.uint8, device=self.device)
            self.ctx.domain_score(rows, n, self.disc_cols[t], dom_t, ccols, dom_c, cooc,
                                  self.torch.from_numpy(hist_t).to(self.device), taus, self.n_rows_global, beta,
                                  top1, prob, weak)
            self.ctx.bitmap_clear_rows(res.bitmaps[t], rows, weak, n)
            removed += int(weak.sum().item())
            res.last_domain = (rows, top1, prob, weak)
        res.weak_removed = removed
        return removed

    # ---- ErrorModel.detect -----------------------------------------------------------------------
    def detect(self, detectors, targets, discrete_thres, opts, given_cells=None):
        """detectors: list of dicts {"type": null|domain|regex|constraint|outlier, ...}
        given_cells: optional (positions, attrs) supplied by setErrorCells."""
        res = DetectResult()
        continuous = self.table.continuous_attrs
        target_attrs = [a for a in self.table.names if not targets or a in set(targets)]
        res.domain_stats = self.discretize(discrete_thres)
        res.disc_attrs = list(self.disc_cols.keys())
        bitmaps, fused = {}, {}
        if given_cells is not None:
            bitmaps = self.bitmaps_from_cells(*given_cells)
        else:
            for det in detectors:
                own = det.get("targets") or []
                tg = [a for a in target_attrs if a in set(own)] if own else target_attrs
                kind = det["type"]
                if kind == "null":
                    self.detect_null(tg, bitmaps, fused)
                elif kind == "regex":
                    self.detect_regex(det["attr"], det["regex"], tg, bitmaps)
                elif kind == "domain":
                    if det["attr"] in tg:
                        rx = self.domain_values_regex(det["attr"], det.get("values", []), det.get("autofill", False),
                                                      det.get("min_count_thres", 12))
                        if rx is not None:
                            self.detect_regex(det["attr"], rx, tg, bitmaps)
                elif kind == "constraint":
                    self.detect_constraints(det.get("path", ""), det.get("constraints", ""), tg, bitmaps)
                elif kind == "outlier":
                    self.detect_outliers(tg, bitmaps, det.get("approx", False))
                else:
                    raise ValueError("unknown detector type: {}".format(kind))
        # one fused pass: NULL bits of the discretised targets + every histogram
        self.scan_hist([a for a in res.disc_attrs if a not in self._hist_cache], fused)
        res.bitmaps = bitmaps
        res.n_cells = dict(zip(bitmaps.keys(), self.ctx.bitmap_count_many(list(bitmaps.values()), self.n_rows)))
        res.n_cells_detected = dict(res.n_cells)     # before the weak-label pruning of the domain analysis
        total = sum(res.n_cells.values())
        if self.dist is not None:
            t = self.torch.tensor([total], dtype=self.torch.int64, device=self.device)
            self.dist.sum_(t)
            total = int(t.item())
        if total == 0:
            res.domain_stats = {}
            return res
        local_noisy = self.torch.tensor([1 if res.n_cells.get(a, 0) > 0 else 0 for a in self.table.names],
                                        dtype=self.torch.int32, device=self.device)
        if self.dist is not None:
            self.dist.max_(local_noisy)
        flags = local_noisy.cpu().numpy()
        res.noisy_columns = [a for a, f in zip(self.table.names, flags) if f]
        if len(res.disc_attrs) == 0:
            res.domain_stats = {}
            return res
        res.target_columns = [a for a in res.noisy_columns if a in self.disc_cols]
        if len(res.target_columns) == 0 or len(res.disc_attrs) <= 1:
            return res
        stats, tables, having = self.compute_attr_stats(
            res.target_columns, res.domain_stats, opts["error.attr_freq_ratio_threshold"],
            opts["error.pairwise_freq_ratio_threshold"], opts["error.max_attrs_to_compute_pairwise_stats"])
        res.pairwise_stats = stats
        if given_cells is None:
            self.prune_weak_labels(res, tables, having, continuous, opts["error.max_attrs_to_compute_domains"],
                                   opts["error.domain_threshold_alpha"], opts["error.domain_threshold_beta"])
            res.n_cells = dict(zip(res.bitmaps.keys(),
                                   self.ctx.bitmap_count_many(list(res.bitmaps.values()), self.n_rows)))
        return res

    # ---- cell frames -----------------------------------------------------------------------------
    def cells_of(self, res, attrs=None):
        """-> list of (attr, row positions np.int64, current codes np.int32) for the error cells."""
        out = []
        for a in self.table.names:
            if a not in res.bitmaps or (attrs is not None and a not in attrs):
                continue
            rows = self.bitmap_rows(res.bitmaps[a])
            n = int(rows.numel())
            if n == 0:
                continue
            cur = self.torch.empty(n, dtype=self.torch.int32, device=self.device)
            self.ctx.gather(self.dt.col(a), rows, n, cur)
            out.append((a, rows.cpu().numpy().astype(np.int64), cur.cpu().numpy()))
        return out

    # ---- repair ----------------------------------------------------------------------------------
    def build_dirty_tile(self, res, target_columns):
        """a10: dirty rows (rows with >= 1 error cell in a target column) gathered row-major with the
        error cells masked to NULL.  -> (dirty_rows, tile int32 [D, K], ctile float64 [D, Kc])"""
        rowmask = self.new_bitmap()
        for a in target_columns:
            if a in res.bitmaps:
                self.ctx.bitmap_or(rowmask, res.bitmaps[a], self.n_rows)
        drows = self.bitmap_rows(rowmask)
        D = int(drows.numel())
        K = len(self.table.columns)
        masks = [res.bitmaps.get(c.name) if c.name in target_columns else None for c in self.table.columns]
        tile = self.torch.empty((max(D, 1), K), dtype=self.torch.int32, device=self.device)
        # NULL bitmap of every tile column, taken while gathering (the chain's work lists)
        self.tile_nulls = self.torch.zeros((K, (D + 31) // 32 + 1), dtype=self.torch.int32, device=self.device)
        self.ctx.gather_rows_masked([self.dt.col(c.name) for c in self.table.columns], masks, drows, D, tile,
                                    null_out=self.tile_nulls)
        cont = [c for c in self.table.columns if c.continuous]
        ctile = None
        if cont:
            ctile = self.torch.empty((max(D, 1), len(cont)), dtype=self.torch.float64, device=self.device)
            cmasks = [res.bitmaps.get(c.name) if c.name in target_columns else None for c in cont]
            self.ctx.gather_rows_masked([self.dt.val(c.name) for c in cont], cmasks, drows, D, ctile, f64=True)
        return drows, tile, ctile

    def sample_rows_masked(self, res, target_columns, rows_np):
        """Training sample: the listed rows of the repair base (error cells masked), on the host."""
        rows = self.torch.from_numpy(np.ascontiguousarray(rows_np, dtype=np.int32)).to(self.device)
        n = len(rows_np)
        K = len(self.table.columns)
        masks = [res.bitmaps.get(c.name) if c.name in target_columns else None for c in self.table.columns]
        tile = self.torch.empty((max(n, 1), K), dtype=self.torch.int32, device=self.device)
        self.ctx.gather_rows_masked([self.dt.col(c.name) for c in self.table.columns], masks, rows, n, tile)
        cont = [c for c in self.table.columns if c.continuous]
        vals = None
        if cont:
            ctile = self.torch.empty((max(n, 1), len(cont)), dtype=self.torch.float64, device=self.device)
            cmasks = [res.bitmaps.get(c.name) if c.name in target_columns else None for c in cont]
            self.ctx.gather_rows_masked([self.dt.val(c.name) for c in cont], cmasks, rows, n, ctile, f64=True)
            vals = ctile[:n].cpu().numpy()
        return tile[:n].cpu().numpy(), vals

    def valid_training_rows(self, res, y, max_rows, seed=42):
        """Rows whose y is non-NULL after masking; at most `max_rows`, seeded choice without
        replacement in table order (stand-in for the unseeded df.sample, model.py:755-766)."""
        col = self.table.by_name[y]
        invalid = self.new_bitmap()
        self.ctx.lut_scan(self.dt.col(y), self.n_rows, None, 0, invalid)  # NULL cells
        if y in res.bitmaps:
            self.ctx.bitmap_or(invalid, res.bitmaps[y], self.n_rows)
        valid = self.torch.bitwise_not(invalid)
        n_valid = self.ctx.bitmap_count(valid, self.n_rows)
        if n_valid == 0:
            return np.zeros(0, dtype=np.int64), 0
        rows = self.bitmap_rows(valid)
        if n_valid > max_rows:
            rng = np.random.default_rng(seed)
            idx = np.sort(rng.choice(n_valid, size=max_rows, replace=False))
            rows = rows[self.torch.from_numpy(idx).to(self.device)]
        del col
        return rows.cpu().numpy().astype(np.int64), n_valid

    def reset(self):
        """Forget per-run state so that the same resident table can be processed again."""
        self._hist_cache = {}
        self.disc_cols, self.disc_dom = {}, {}

    def close(self):
        self.ctx.close()
