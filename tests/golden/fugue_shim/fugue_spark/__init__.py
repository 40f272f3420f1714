"""Placeholder for `fugue_spark` (imported unconditionally by the reference's fugue.py:10-11);
the Spark engine is never instantiated by the golden generator."""


class SparkExecutionEngine(object):
    def __init__(self, *a, **k):
        raise RuntimeError("SparkExecutionEngine is not available in the golden-generation shim")


class SparkDataFrame(object):
    def __init__(self, *a, **k):
        raise RuntimeError("SparkDataFrame is not available in the golden-generation shim")
