"""Factor graphs: the model / posterior containers and the topological log-density and sampling walks
(mxfusion/models/factor_graph.py:31-643; the walks are :192-297).

Own design, not the reference's networkx graph: nodes keep their own edge lists, a graph keeps the
nodes the user anchored on it (``m.x = ...``, ``q[v]``) and derives its component set as the ancestor
closure of those anchors, cached against a global edge-version counter."""
from collections import OrderedDict
from uuid import uuid4
import warnings

import torch

from ..components.model_component import ModelComponent
from ..components.variables.variable import Variable, VariableType
from ..components.variables.runtime_variable import expectation
from ..components.factor import Factor
from ..components.functions.mxfusion_function import FunctionEvaluation, MXFusionFunction
from ..components.distributions.distribution import Distribution
from ..common.exceptions import ModelSpecificationError, InferenceError


class FactorGraph(object):
    def __init__(self, name=None, verbose=False):
        d = self.__dict__
        d['name'] = name
        d['_uuid'] = str(uuid4()).replace('-', '_')
        d['_anchors'] = OrderedDict()      # uuid -> component
        d['_var_ties'] = {}
        d['_verbose'] = verbose
        d['_cache'] = None

    def __repr__(self):
        return '\n'.join(str(f) for f in self.ordered_factors)

    # registration ----------------------------------------------------------------------------------
    def __setattr__(self, name, value):
        if isinstance(value, ModelComponent):
            if self._verbose:
                print("Variable " + name + " = " + str(value))
            value.name = name
            value.graph = self
            self._anchors[value.uuid] = value
            self.__dict__['_cache'] = None
        self.__dict__[name] = value

    def _closure(self):
        """Ancestor closure of the anchors: (components by uuid, factors in topological order)."""
        key = (len(self._anchors), tuple(id(c) for c in self._anchors.values()))
        comps, order, state = OrderedDict(), [], {}

        def visit(node):
            st = state.get(node.uuid)
            if st == 2:
                return
            if st == 1:
                raise ModelSpecificationError("The factor graph has a cycle through " + str(node) + ".")
            state[node.uuid] = 1
            for a in node.attributes:
                visit(a)
            for _, pred in node._in:
                visit(pred)
            state[node.uuid] = 2
            comps[node.uuid] = node
            if isinstance(node, Factor):
                order.append(node)
                for _, out in node._out:         # sibling outputs of a visited factor belong to the graph
                    if out.uuid not in state:
                        state[out.uuid] = 2
                        comps[out.uuid] = out
        for c in list(self._anchors.values()):
            visit(c)
        return comps, order

    @property
    def components(self):
        return self._closure()[0]

    @property
    def variables(self):
        return OrderedDict((u, c) for u, c in self._closure()[0].items() if isinstance(c, Variable))

    @property
    def ordered_factors(self):
        return self._closure()[1]

    @property
    def distributions(self):
        return {f.uuid: f for f in self.ordered_factors if isinstance(f, Distribution)}

    @property
    def functions(self):
        return {f.uuid: f for f in self.ordered_factors if isinstance(f, FunctionEvaluation)}

    @property
    def modules(self):
        from ..modules.module import Module
        return OrderedDict((f.uuid, f) for f in self.ordered_factors if isinstance(f, Module))

    @property
    def var_ties(self):
        return self._var_ties

    def __contains__(self, key):
        key = key.uuid if isinstance(key, ModelComponent) else key
        return key in self._closure()[0]

    def __getitem__(self, key):
        key = key.uuid if isinstance(key, ModelComponent) else key
        return self._closure()[0][key]

    def get_parameters(self, excluded=None, include_inherited=True):
        """PARAMETER variables not in `excluded` (factor_graph.py:360-378)."""
        excluded = set() if excluded is None else set(excluded)
        return [v for v in self.variables.values()
                if v.type == VariableType.PARAMETER and v.uuid not in excluded and
                (include_inherited or not v.isInherited)]

    def get_constants(self):
        return [v for v in self.variables.values() if v.type == VariableType.CONSTANT]

    def get_latent_variables(self, observed):
        obs = set(v.uuid if isinstance(v, ModelComponent) else v for v in observed)
        return [v for v in self.variables.values() if v.type == VariableType.RANDVAR and v.uuid not in obs]

    # the two walks ---------------------------------------------------------------------------------
    def log_pdf(self, F, variables, targets=None):
        """factor_graph.py:192-238: sum over factors of F.sum(mean over samples(log_pdf))."""
        from ..modules.module import Module
        if targets is not None:
            targets = set(t.uuid if isinstance(t, ModelComponent) else t for t in targets)
        logL = 0.
        batch = []
        for f in self.ordered_factors:
            if isinstance(f, FunctionEvaluation):
                outcome = f.eval(F=F, variables=variables, always_return_tuple=True)
                for v, (_, var) in zip(outcome, f.outputs):
                    if var.uuid in variables:
                        warnings.warn('Function evaluation in FactorGraph.log_pdf: the outcome variable ' +
                                      str(var.uuid) + ' of ' + str(f) + ' has already existed in the variable set.')
                    variables[var.uuid] = v
            elif isinstance(f, Distribution):
                if targets is None or f.random_variable.uuid in targets:
                    ops_ = f.log_pdf_operands(F, variables) if hasattr(f, 'log_pdf_operands') else None
                    if ops_ is not None:
                        batch.append(ops_)          # Normal factors: evaluated together below, one launch each way
                        continue
                    term = f.log_pdf_sum(F=F, variables=variables)
                    logL = term if (isinstance(logL, float) and logL == 0.) else logL + term
            elif isinstance(f, Module):
                if targets is None:
                    module_targets = [v.uuid for _, v in f.outputs if v.uuid in variables]
                else:
                    module_targets = [v.uuid for _, v in f.outputs if v.uuid in targets]
                if len(module_targets) > 0:
                    e = expectation(F, f.log_pdf(F=F, variables=variables, targets=module_targets))
                    term = e.reshape(()) if e.numel() == 1 else torch.sum(e)     # F.sum of one element: a view
                    logL = term if (isinstance(logL, float) and logL == 0.) else logL + term
            else:
                raise ModelSpecificationError("There is an object in the factor graph that isn't a factor.")
        if batch:
            from .. import ops
            term = ops.normal_log_pdf_sum_multi(batch).reshape(())
            logL = term if (isinstance(logL, float) and logL == 0.) else logL + term
        return logL

    def draw_samples(self, F, variables, num_samples=1, targets=None):
        """factor_graph.py:240-297."""
        from ..modules.module import Module
        samples = {}
        pending = []            # independent Normal draws, issued together (one launch, one adjoint launch)

        def flush():
            if not pending:
                return
            from .. import ops
            from ..components.distributions.random_gen import step_counter
            seed = pending[0][1][3][0]
            dev = pending[0][1][0].device
            ws = ops.normal_draw_multi([(m, v, ns) for _, (m, v, ns, _) in pending], seed,
                                       [st[1] for _, (_, _, _, st) in pending], step_counter(dev))
            for (fac, _), w in zip(pending, ws):
                var = fac.outputs[0][1]
                variables[var.uuid] = w
                samples[var.uuid] = w
            del pending[:]

        for f in self.ordered_factors:
            if pending and any(v.uuid not in variables for _, v in f.inputs):
                flush()                                    # this factor consumes a draw that is still pending
            if isinstance(f, FunctionEvaluation):
                outcome = f.eval(F=F, variables=variables, always_return_tuple=True)
                for v, (_, var) in zip(outcome, f.outputs):
                    variables[var.uuid] = v
                    samples[var.uuid] = v
            elif isinstance(f, Distribution):
                known = [v.uuid in variables for _, v in f.outputs]
                if all(known):
                    continue
                elif any(known):
                    raise InferenceError("Part of the outputs of the distribution " + f.__class__.__name__ +
                                         " has been observed!")
                if hasattr(f, 'draw_operands') and all(v.uuid in variables for _, v in f.inputs):
                    ops_ = f.draw_operands(F, variables, num_samples)
                    if ops_ is not None:
                        pending.append((f, ops_))
                        continue
                outcome = f.draw_samples(F=F, num_samples=num_samples, variables=variables,
                                         always_return_tuple=True)
                for v, (_, var) in zip(outcome, f.outputs):
                    variables[var.uuid] = v
                    samples[var.uuid] = v
            elif isinstance(f, Module):
                outcome_uuid = [v.uuid for _, v in f.outputs]
                outcome = f.draw_samples(F=F, variables=variables, num_samples=num_samples, targets=outcome_uuid)
                for v, uuid in zip(outcome, outcome_uuid):
                    variables[uuid] = v
                    samples[uuid] = v
            else:
                raise ModelSpecificationError("There is an object in the factor graph that isn't a factor.")
        flush()
        if targets:
            return tuple(samples[t.uuid if isinstance(t, ModelComponent) else t] for t in targets)
        return samples

    def as_json(self):
        comps = self.components
        return {'name': self.name, 'uuid': self._uuid, 'class': type(self).__name__,
                'components': [c.as_json() for c in comps.values()],
                'edges': [[src.uuid, dst.uuid, n] for dst in comps.values() for n, src in dst._in]}


def _anchor(graph, component, name=None):
    """Register `component` on `graph` without going through attribute assignment."""
    if name is not None:
        setattr(graph, name, component)
    else:
        component.graph = graph
        graph._anchors[component.uuid] = component
    return component


FactorGraph.add_component = _anchor
