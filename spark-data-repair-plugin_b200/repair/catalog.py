"""A tiny in-process table catalog so that ``setTableName("adult")`` / ``setInput("adult")`` keep
working without a Spark session: ``repair.catalog.register("adult", dataframe)``.  If a SparkSession
is active, unknown names are looked up there and collected."""
from .utils import AnalysisException

_tables = {}


def register(name, df):
    _tables[name] = df
    return df


def unregister(name):
    _tables.pop(name, None)


def table(name):
    if name in _tables:
        return _tables[name]
    short = name.split(".")[-1]
    if short in _tables:
        return _tables[short]
    try:  # pragma: no cover - pyspark is not installed in the build image
        from pyspark.sql import SparkSession  # type: ignore
        spark = SparkSession.getActiveSession()
        if spark is not None:
            return spark.table(name).toPandas()
    except ImportError:
        pass
    raise AnalysisException("Table or view not found: {}".format(name))
