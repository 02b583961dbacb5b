"""RandomVariableDifferentiableAAD over RandomVariableCuda — composability with backward-mode algorithmic differentiation.

The reference states the contract in README.md:50-52 ("RandomVariableCudaFactory can be combined with AAD wrappers, for
example RandomVariableDifferentiableAAD ... objects allowing for AAD have higher priority, AAD on GPU has higher priority
than AAD on CPU") and lists `RandomVariableDifferentiableAADFactory` among the interchangeable factories (README.md:117-119).
The wrapper class itself lives in finmath-lib (net.finmath.montecarlo.automaticdifferentiation.backward, not vendored in
the reference tree), so this is a restatement of its published behaviour, not of reference source:
  * a differentiable variable wraps the values of ANY inner RandomVariable type and a node of an operator tree;
  * every operation computes its values with the inner type's own method (so on RandomVariableCuda they are recorded on
    the op-tape and fused like any other chain) and records (operator, arguments);
  * getGradient() is one reverse sweep over the tree; the partial derivatives are again expressed only through inner-type
    operations (mult, add, sub, div, squared, exp, log, sqrt, invert, choose, average ...), i.e. they run on the GPU;
  * type priority = AAD_PRIORITY_OFFSET + inner priority, so a differentiable operand takes over from a plain one
    (RandomVariableCuda.java:1392-1395 pattern) and "AAD on GPU" (offset + 20) outranks "AAD on CPU" (offset + 1).
Retention policy (the reference only states the contract, README.md:50-52,119). A node of the operator tree
keeps exactly the VALUES its own derivative rule reads and nothing else: add / sub / average / a product with a scalar keep
none, mult / div keep the two operands, sqrt / exp / invert their own result, log / squared / abs / pow the argument, accrue /
discount value and rate, choose the trigger. On the GPU type this decides what is ever MATERIALISED: a primal intermediate that no
rule needs has no handle left once the caller's temporary dies, so the op-tape fuses it away instead of storing a vector per
node (RETENTION = "all" restores the store-everything behaviour for comparison: benchmarks/aad_footprint.py). During the
reverse sweep an adjoint is dropped as soon as it has been propagated, and getGradient(release=True) also drops each node's
retained values right after its rule has run, so the footprint shrinks as the sweep proceeds (the tree is then single-use).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from .random_variable import RandomVariable, RandomVariableCuda, RandomVariableCudaFactory

AAD_PRIORITY_OFFSET = 1000
_ids = itertools.count(1)


RETENTION = "needed"          # "needed": a node keeps only what its derivative rule reads; "all": every node also keeps its own values

_KEEPS_OWN_VALUE = {"sqrt", "exp", "invert"}


class _Node:
    __slots__ = ("id", "op", "args", "values", "param")

    def __init__(self, op: Optional[str], args: Sequence["_Node" | None], values: RandomVariable, param=None):
        self.id = next(_ids)
        self.op, self.args, self.param = op, list(args), param
        self.values = values if (RETENTION == "all" or op in _KEEPS_OWN_VALUE) else None


class RandomVariableDifferentiableAAD(RandomVariable):
    """finmath-lib's RandomVariableDifferentiableAAD restated over an arbitrary inner RandomVariable type."""

    def __init__(self, values: RandomVariable, node: Optional[_Node] = None):
        self.values = values
        self.node = node if node is not None else _Node(None, [], values)

    # ------------------------------------------------------------------ interface plumbing
    def getID(self) -> int: return self.node.id
    def getValues(self) -> RandomVariable: return self.values
    def getTypePriority(self) -> int: return AAD_PRIORITY_OFFSET + self.values.getTypePriority()
    def getFiltrationTime(self) -> float: return self.values.getFiltrationTime()
    def isDeterministic(self) -> bool: return self.values.isDeterministic()
    def size(self) -> int: return self.values.size()
    def get(self, i: int) -> float: return self.values.get(i)
    def getRealizations(self): return self.values.getRealizations()
    def doubleValue(self) -> float: return self.values.doubleValue()
    def getAverage(self, *a) -> float: return self.values.getAverage(*a)
    def getVariance(self, *a) -> float: return self.values.getVariance(*a)
    def getStandardError(self, *a) -> float: return self.values.getStandardError(*a)
    def getMin(self) -> float: return self.values.getMin()
    def getMax(self) -> float: return self.values.getMax()

    @staticmethod
    def _val(x):
        return x.values if isinstance(x, RandomVariableDifferentiableAAD) else x

    @staticmethod
    def _arg(x):
        return x.node if isinstance(x, RandomVariableDifferentiableAAD) else None

    def _new(self, op: str, values: RandomVariable, args, param=None) -> "RandomVariableDifferentiableAAD":
        return RandomVariableDifferentiableAAD(values, _Node(op, [self._arg(a) for a in args], values, param))

    # ------------------------------------------------------------------ operations (values through the inner type)
    def add(self, x): return self._new("add", self.values.add(self._val(x)), [self, x])
    def sub(self, x): return self._new("sub", self.values.sub(self._val(x)), [self, x])
    def bus(self, x): return self._new("sub", self.values.bus(self._val(x)), [x, self])
    def mult(self, x):
        # a factor is kept only if the OTHER operand is differentiable (d/dx (x y) = y): x * scalar keeps one double
        keep_self = self.values if isinstance(x, RandomVariableDifferentiableAAD) else None
        return self._new("mult", self.values.mult(self._val(x)), [self, x], param=("operands", keep_self, self._val(x)))
    def div(self, x): return self._new("div", self.values.div(self._val(x)), [self, x], param=("operands", self.values, self._val(x)))
    def vid(self, x): return self._new("div", self.values.vid(self._val(x)), [x, self], param=("operands", self._val(x), self.values))
    def squared(self): return self._new("squared", self.values.squared(), [self], param=self.values)
    def sqrt(self): return self._new("sqrt", self.values.sqrt(), [self])
    def exp(self): return self._new("exp", self.values.exp(), [self])
    def log(self): return self._new("log", self.values.log(), [self], param=self.values)
    def invert(self): return self._new("invert", self.values.invert(), [self])
    def abs(self): return self._new("abs", self.values.abs(), [self], param=self.values)
    def pow(self, e: float): return self._new("pow", self.values.pow(e), [self], param=(self.values, float(e)))
    def cap(self, c): return self._new("cap", self.values.cap(self._val(c)), [self, c], param=("operands", self.values, self._val(c)))
    def floor(self, c): return self._new("floor", self.values.floor(self._val(c)), [self, c], param=("operands", self.values, self._val(c)))
    def average(self): return self._new("average", self.values.average(), [self])

    def accrue(self, rate, p: float):
        return self._new("accrue", self.values.accrue(self._val(rate), p), [self, rate], param=(self.values, self._val(rate), float(p)))

    def discount(self, rate, p: float):
        return self._new("discount", self.values.discount(self._val(rate), p), [self, rate], param=(self.values, self._val(rate), float(p)))

    def addProduct(self, f1, f2):
        return self._new("addProduct", self.values.addProduct(self._val(f1), self._val(f2)), [self, f1, f2], param=(self._val(f1), self._val(f2)))

    def choose(self, a, b):
        return self._new("choose", self.values.choose(self._val(a), self._val(b)), [self, a, b], param=self.values)

    # ------------------------------------------------------------------ reverse sweep
    def getGradient(self, independents: Optional[Sequence["RandomVariableDifferentiableAAD"]] = None, release: bool = False) -> Dict[int, RandomVariable]:
        """d(self)/d(leaf) for every leaf of the operator tree (or the given variables), as inner-type random variables.
        release=True: every node drops the values it retained as soon as its derivative rule has run (single-use tree)."""
        one = _scalar_like(self.values, 1.0)
        adj: Dict[int, RandomVariable] = {self.node.id: one}
        # nodes reachable from self, processed in decreasing id (ids increase along every edge)
        seen, stack, order = {self.node.id}, [self.node], []
        while stack:
            nd = stack.pop()
            order.append(nd)
            for a in nd.args:
                if a is not None and a.id not in seen:
                    seen.add(a.id); stack.append(a)
        order.sort(key=lambda nd: -nd.id)
        want = None if independents is None else {v.getID() for v in independents}
        grad: Dict[int, RandomVariable] = {}
        for nd in order:
            a = adj.pop(nd.id, None)
            if a is None:
                continue
            if nd.op is None:
                if want is None or nd.id in want:
                    grad[nd.id] = a
                continue
            if want is not None and nd.id in want:
                grad[nd.id] = a
            for k, arg in enumerate(nd.args):
                if arg is None:
                    continue
                contrib = _propagate(nd, k, a)
                if contrib is None:
                    continue
                adj[arg.id] = adj[arg.id].add(contrib) if arg.id in adj else contrib
            if release:
                nd.param = None; nd.values = None
        return grad


def _scalar_like(v: RandomVariable, x: float) -> RandomVariable:
    return type(v)(x)                  # the inner type's own (value) constructor: RandomVariableCuda(x), the CPU twin's, ...


def _indicator(trigger: RandomVariable, if_nonneg: float, if_neg: float) -> RandomVariable:
    return trigger.choose(_scalar_like(trigger, if_nonneg), _scalar_like(trigger, if_neg))


def _propagate(nd: _Node, k: int, a: RandomVariable) -> Optional[RandomVariable]:
    """adjoint contribution of node nd to its k-th argument: a * d(nd)/d(arg_k), inner-type operations only."""
    op, v, p = nd.op, nd.values, nd.param
    if op == "add": return a
    if op == "sub": return a if k == 0 else a.mult(-1.0)
    if op == "mult": return a.mult(p[2] if k == 0 else p[1])
    if op == "div":                                        # x / y
        x, y = p[1], p[2]
        return a.div(y) if k == 0 else a.mult(x).div(y.squared()).mult(-1.0)
    if op == "squared": return a.mult(p).mult(2.0)
    if op == "sqrt": return a.div(v).mult(0.5)
    if op == "exp": return a.mult(v)
    if op == "log": return a.div(p)
    if op == "invert": return a.mult(v.squared()).mult(-1.0)
    if op == "abs": return a.mult(_indicator(p, 1.0, -1.0))
    if op == "pow": return a.mult(p[0].pow(p[1] - 1.0)).mult(p[1])
    if op == "cap":                                        # min(x, c): d/dx = 1{x < c}
        x, c = p[1], p[2]
        ind = _indicator(_as_rv(c, x).sub(x), 1.0, 0.0)
        return a.mult(ind) if k == 0 else a.mult(ind.bus(1.0))
    if op == "floor":                                      # max(x, c): d/dx = 1{x > c}
        x, c = p[1], p[2]
        ind = _indicator(x.sub(_as_rv(c, x)), 1.0, 0.0)
        return a.mult(ind) if k == 0 else a.mult(ind.bus(1.0))
    if op == "average": return a.average() if not a.isDeterministic() else a
    if op == "accrue":                                     # x * (1 + r p)
        x, r, per = p
        return a.mult(_as_rv(r, x).mult(per).add(1.0)) if k == 0 else a.mult(x).mult(per)
    if op == "discount":                                   # x / (1 + r p)
        x, r, per = p
        den = _as_rv(r, x).mult(per).add(1.0)
        return a.div(den) if k == 0 else a.mult(x).mult(-per).div(den.squared())
    if op == "addProduct":                                 # x + f1 f2
        f1, f2 = p
        return a if k == 0 else a.mult(f2 if k == 1 else f1)
    if op == "choose":                                     # trigger >= 0 ? a : b  (no sensitivity to the trigger itself)
        if k == 0: return None
        return a.mult(_indicator(p, 1.0, 0.0) if k == 1 else _indicator(p, 0.0, 1.0))
    raise NotImplementedError(op)


def _as_rv(x, like: RandomVariable) -> RandomVariable:
    return x if isinstance(x, RandomVariable) else _scalar_like(like, float(x))


class RandomVariableDifferentiableAADFactory:
    """RandomVariableDifferentiableAADFactory(innerFactory): createRandomVariable returns differentiable variables whose
    values are created by the inner factory (README.md:117-119)."""

    def __init__(self, randomVariableFactoryForNonDifferentiable=None):
        self.inner = randomVariableFactoryForNonDifferentiable or RandomVariableCudaFactory()

    def createRandomVariable(self, *args) -> RandomVariableDifferentiableAAD:
        return RandomVariableDifferentiableAAD(self.inner.createRandomVariable(*args))

    def createRandomVariableNonDifferentiable(self, *args) -> RandomVariable:
        return self.inner.createRandomVariable(*args)
