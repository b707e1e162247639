"""Stage-graph IR consumed by the runner.

Same vocabulary as the reference's IR (dampr/runner.py:17-135: Source, GMap, GReduce, GSink,
Graph) so the DSL contract is unchanged: graphs are immutable values, every add_* returns a new
graph plus the Source naming the stage's output.
"""
import itertools

from .operators import Mapper, Reducer, Combiner


class Source(object):
    """Handle of a dataset in a graph (an input or a stage output). Identity is a global serial."""
    _serial = itertools.count()

    def __init__(self, name):
        self.name = name
        self.cnt = next(Source._serial)

    def __hash__(self):
        return hash(self.cnt)

    def __eq__(self, other):
        return isinstance(other, Source) and self.cnt == other.cnt

    def __repr__(self):
        return "Source[`%s`]" % self.name

    __str__ = __repr__


class _Stage(object):
    kind = "?"

    def __init__(self, output, inputs, options):
        self.output = output
        self.inputs = list(inputs)
        self.options = dict(options) if options else {}

    def __repr__(self):
        return self.kind


class GMap(_Stage):
    kind = "Map"

    def __init__(self, output, inputs, mapper, combiner=None, shuffler=None, options=None):
        _Stage.__init__(self, output, inputs, options)
        self.mapper = mapper
        self.combiner = combiner
        self.shuffler = shuffler


class GReduce(_Stage):
    kind = "Reducer"

    def __init__(self, output, inputs, reducer, options=None):
        _Stage.__init__(self, output, inputs, options)
        self.reducer = reducer


class GSink(_Stage):
    kind = "Sink"

    def __init__(self, output, inputs, mapper, path, options=None):
        _Stage.__init__(self, output, inputs, options)
        self.mapper = mapper
        self.path = path

    def __repr__(self):
        return "Sink[path=%s]" % self.path


class Graph(object):
    def __init__(self, inputs=None, stages=None):
        self.inputs = dict(inputs) if inputs else {}
        self.stages = list(stages) if stages else []

    def _clone(self):
        return Graph(self.inputs, self.stages)

    def add_input(self, dataset):
        g = self._clone()
        src = Source("Input:%d" % len(self.inputs))
        g.inputs[src] = dataset
        return src, g

    def _add(self, stage_cls, label, default, inputs, *args, **kw):
        assert all(isinstance(i, Source) for i in inputs)
        name = kw.pop("name", None) or default
        src = Source(name.format(label))
        g = self._clone()
        g.stages.append(stage_cls(src, inputs, *args, **kw))
        return src, g

    def add_mapper(self, inputs, mapper, combiner=None, shuffler=None, name=None, options=None):
        assert isinstance(mapper, Mapper)
        assert combiner is None or isinstance(combiner, Combiner)
        return self._add(GMap, len(self.stages), "Map: {}", inputs, mapper, combiner, shuffler,
                         name=name, options=options)

    def add_reducer(self, inputs, reducer, name=None, options=None):
        assert isinstance(reducer, Reducer)
        return self._add(GReduce, len(self.stages), "Reduce: {}", inputs, reducer, name=name, options=options)

    def add_sink(self, inputs, mapper, path, name=None, options=None):
        assert isinstance(mapper, Mapper)
        return self._add(GSink, path, "Sink: {}", inputs, mapper, path, name=name, options=options)

    def union(self, other):
        g = self._clone()
        g.inputs.update(other.inputs)
        have = set(id(s) for s in g.stages)
        for s in other.stages:
            if id(s) not in have:
                g.stages.append(s)
                have.add(id(s))
        return g
