"""``delphi.misc`` (python/repair/misc.py:27-131): of the reference's helper functions only the one on
the repair path is provided -- ``repair()``, which applies a frame of predicted updates
``(row_id, attribute, repaired)`` to an input table (RepairMiscApi.scala:184-247)."""
from typing import Dict, List

import numpy as np
import pandas as pd

from . import catalog
from .utils import AnalysisException, argtype_check, row_positions


def _to_double(v):
    """Spark's CAST(string AS DOUBLE): also takes a trailing d/D/f/F type suffix; unparsable -> NULL."""
    if isinstance(v, str):
        v = v.strip()
        if v[-1:] in "dDfF" and v[:-1]:
            v = v[:-1]
    try:
        return float(v)
    except (TypeError, ValueError):
        return float("nan")


class RepairMisc():

    def __init__(self) -> None:
        self.opts: Dict[str, str] = {}

    @argtype_check
    def option(self, key: str, value: str) -> "RepairMisc":
        self.opts[str(key)] = str(value)
        return self

    @argtype_check
    def options(self, options: Dict[str, str]) -> "RepairMisc":
        self.opts.update(options)
        return self

    def _check_required_options(self, required: List[str]) -> None:
        if not all(opt in self.opts.keys() for opt in required):
            raise ValueError("Required options not found: {}".format(", ".join(required)))

    def repair(self) -> pd.DataFrame:
        """Applies predicted repair updates into an input table: a cell whose (row id, attribute) is
        listed takes the `repaired` value -- cast to the column's type, rounded first for integral
        columns (RepairMiscApi.scala:224-230) -- every other cell is kept; when a cell is listed twice
        the last update wins (Spark's map_from_entries keeps the last duplicate key under
        spark.sql.mapKeyDedupPolicy=LAST_WIN; the default policy raises instead)."""
        self._check_required_options(["repair_updates", "table_name", "row_id"])
        row_id = self.opts["row_id"]
        name = self.opts["table_name"]
        if self.opts.get("db_name"):
            name = "{}.{}".format(self.opts["db_name"], name)
        table = catalog.table(name)
        updates = catalog.table(self.opts["repair_updates"])
        if not all(c in updates.columns for c in (row_id, "attribute", "repaired")):
            raise AnalysisException("Table '{}' must have '{}', 'attribute', and 'repaired' columns".format(
                self.opts["repair_updates"], row_id))
        out = table.copy()
        ids = out[row_id].to_numpy()
        for attr, grp in updates.groupby("attribute", sort=False):
            if attr not in out.columns or attr == row_id:
                continue
            col = out[attr]
            # vectorised join on the row id (no Python dict over the table's rows)
            pos, found = row_positions(ids, grp[row_id].to_numpy())
            rows = pos[found].tolist()
            vals = [None if v is None or (isinstance(v, float) and v != v) else v
                    for v, f in zip(grp["repaired"].tolist(), found.tolist()) if f]
            if not rows:
                continue
            if pd.api.types.is_integer_dtype(col.dtype):
                arr = col.astype("Int64").to_numpy(dtype=object, copy=True)
                for i, v in zip(rows, vals):
                    d = np.nan if v is None else _to_double(v)
                    # Spark's round() is half-up (away from zero), unlike numpy's half-even
                    arr[i] = pd.NA if d != d else int(np.floor(abs(d) + 0.5) * (1 if d >= 0 else -1))
                out[attr] = pd.array(arr, dtype="Int64")
            elif pd.api.types.is_float_dtype(col.dtype):
                arr = col.to_numpy(dtype=np.float64, copy=True)
                for i, v in zip(rows, vals):
                    arr[i] = np.nan if v is None else _to_double(v)
                out[attr] = arr
            else:
                arr = col.to_numpy(dtype=object, copy=True)
                for i, v in zip(rows, vals):
                    arr[i] = v
                out[attr] = arr
        return out
