"""Minimal stand-in for the slice of Fugue 0.4.7 that the reference's ``node2vec/fugue.py``
uses, on plain pandas -- GOLDEN-GENERATION TOOLING ONLY (tests/golden/make_golden.py).

It lets the UNMODIFIED reference ``trim_index`` / ``random_walk`` run in this container
(Fugue itself is not installable offline) so their outputs can be recorded as fixtures.
Semantics follow Fugue's NativeExecutionEngine: a workflow is evaluated eagerly on one pandas
partition; ``partition(by=...)`` groups by key in ascending key order (``presort`` sorts rows
inside a group, stable); ``transform`` feeds a function either a pandas frame or an iterable of
dict rows according to its first parameter's annotation and collects the yielded dicts; joins
are natural joins on the common columns keeping the left frame's row order; NaN / missing
values reach dict rows as ``None``.  Output schemas come from the ``# schema:`` hint comment
above the transformer (as Fugue reads them) or from ``schema="*"``.
"""
import inspect
import re
import typing
from typing import Any, Dict, Iterable, List, Optional

import pandas as pd

__all__ = ["FugueWorkflow", "DataFrame", "ExecutionEngine", "PandasDataFrame", "ArrayDataFrame",
           "NativeExecutionEngine"]


class Schema(list):
    @property
    def names(self) -> List[str]:
        return list(self)


class DataFrame(object):
    """Concrete frame: a pandas DataFrame with a schema."""

    def __init__(self, df: Any = None, schema: Any = None):
        if isinstance(df, DataFrame):
            df = df.as_pandas()
        if not isinstance(df, pd.DataFrame):
            cols = [c.split(":")[0] for c in schema.split(",")] if isinstance(schema, str) else list(schema)
            df = pd.DataFrame(list(df), columns=cols)
        self._df = df.reset_index(drop=True)

    @property
    def schema(self) -> Schema:
        return Schema(self._df.columns)

    @property
    def native(self) -> pd.DataFrame:
        return self._df

    def as_pandas(self) -> pd.DataFrame:
        return self._df

    def count(self) -> int:
        return len(self._df)

    def rename(self, columns: Dict[str, str]) -> "DataFrame":
        return type(self)(self._df.rename(columns=columns))

    def __getitem__(self, cols: List[str]) -> "DataFrame":
        return type(self)(self._df[list(cols)])


class PandasDataFrame(DataFrame):
    pass


class ArrayDataFrame(DataFrame):
    pass


class ExecutionEngine(object):
    def __init__(self, conf: Optional[Dict[str, Any]] = None):
        self.conf: Dict[str, Any] = dict(conf or {})


class NativeExecutionEngine(ExecutionEngine):
    pass


def _schema_hint(fn) -> Optional[List[str]]:
    """Column names from the `# schema: a:int,b:str` comment right above a transformer."""
    try:
        lines, start = inspect.getsourcelines(fn)
        module_lines = inspect.getsource(inspect.getmodule(fn)).splitlines()
    except (OSError, TypeError):
        return None
    for k in range(start - 2, max(start - 6, -1), -1):
        m = re.match(r"\s*#\s*schema:\s*(.+)$", module_lines[k])
        if m:
            return [c.split(":")[0].strip() for c in m.group(1).split(",")]
        if module_lines[k].strip() and not module_lines[k].strip().startswith("#"):
            break
    return None


def _rows(df: pd.DataFrame) -> Iterable[Dict[str, Any]]:
    cols = list(df.columns)
    for rec in df.itertuples(index=False, name=None):
        yield {c: (None if (v is None or (isinstance(v, float) and v != v)) else v) for c, v in zip(cols, rec)}


class WorkflowDataFrame(object):
    """A node of the (eagerly evaluated) workflow."""

    def __init__(self, df: pd.DataFrame, by: Optional[List[str]] = None, presort: Optional[str] = None):
        self._df, self._by, self._presort = df.reset_index(drop=True), by, presort

    # -- structure ------------------------------------------------------------------
    def partition(self, by: List[str], presort: Optional[str] = None) -> "WorkflowDataFrame":
        return WorkflowDataFrame(self._df, list(by), presort)

    def persist(self) -> "WorkflowDataFrame":
        return self

    def checkpoint(self) -> "WorkflowDataFrame":
        return self

    def compute(self) -> PandasDataFrame:
        return PandasDataFrame(self._df)

    def rename(self, *args: Any, **kwargs: str) -> "WorkflowDataFrame":
        mapping = dict(args[0]) if args else dict(kwargs)
        return WorkflowDataFrame(self._df.rename(columns=mapping))

    def drop(self, cols: List[str]) -> "WorkflowDataFrame":
        return WorkflowDataFrame(self._df.drop(columns=list(cols)))

    def __getitem__(self, cols: List[str]) -> "WorkflowDataFrame":
        return WorkflowDataFrame(self._df[list(cols)])

    # -- joins: natural, on the common columns, left row order kept -------------------------
    def _join(self, other: "WorkflowDataFrame", how: str) -> "WorkflowDataFrame":
        on = [c for c in self._df.columns if c in other._df.columns]
        return WorkflowDataFrame(self._df.merge(other._df, on=on, how=how))

    def inner_join(self, other: "WorkflowDataFrame") -> "WorkflowDataFrame":
        return self._join(other, "inner")

    def left_outer_join(self, other: "WorkflowDataFrame") -> "WorkflowDataFrame":
        return self._join(other, "left")

    # -- transform -----------------------------------------------------------------
    def transform(self, fn, schema: Optional[str] = None, params: Optional[Dict[str, Any]] = None) -> "WorkflowDataFrame":
        params = dict(params or {})
        first = list(inspect.signature(fn).parameters.values())[0]
        wants_frame = first.annotation is pd.DataFrame
        if self._by:
            parts = [p for _, p in self._df.groupby(self._by, sort=True)]
            if self._presort:
                parts = [p.sort_values(self._presort, kind="stable") for p in parts]
        else:
            parts = [self._df]
        out: List[Dict[str, Any]] = []
        for part in parts:
            part = part.reset_index(drop=True)
            arg = part if wants_frame else _rows(part)
            for row in fn(arg, **params):
                out.append(dict(row))
        if schema == "*":
            cols = list(self._df.columns)
        else:
            cols = _schema_hint(fn) or (list(out[0].keys()) if out else list(self._df.columns))
        return WorkflowDataFrame(pd.DataFrame(out, columns=cols))


class FugueWorkflow(object):
    def __init__(self, engine: Optional[ExecutionEngine] = None):
        self.engine = engine

    def df(self, data: Any) -> WorkflowDataFrame:
        if isinstance(data, WorkflowDataFrame):
            return data
        if isinstance(data, DataFrame):
            return WorkflowDataFrame(data.as_pandas())
        return WorkflowDataFrame(pd.DataFrame(data))
