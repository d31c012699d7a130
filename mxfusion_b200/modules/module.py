"""Probabilistic modules: a factor that owns inner factor graphs and the algorithms that compute its
log-density, samples and predictions (mxfusion/modules/module.py:27-479).  This is the plug-in
boundary of the hot path: `attach_log_pdf_algorithms` registers an `InferenceAlgorithm` whose
`compute(F, variables)` the graph walk calls with `log_pdf_scaling` set (module.py:304-322)."""
import warnings

from ..components.factor import Factor
from ..components.variables.variable import Variable, VariableType
from ..components.distributions.random_gen import MXNetRandomGenerator
from ..common.config import get_default_dtype
from ..common.exceptions import ModelSpecificationError


class Module(Factor):
    is_probabilistic = True

    def __init__(self, inputs, outputs, input_names, output_names, rand_gen=None, dtype=None, ctx=None):
        super(Module, self).__init__(inputs=inputs, outputs=outputs, input_names=input_names,
                                     output_names=output_names)
        self._rand_gen = MXNetRandomGenerator if rand_gen is None else rand_gen
        self.dtype = get_default_dtype() if dtype is None else dtype
        self.ctx = ctx
        self._module_graph = None
        self._extra_graphs = []
        self._log_pdf_algorithms = {}
        self._draw_samples_algorithms = {}
        self._prediction_algorithms = {}
        self.log_pdf_scaling = 1

    def __contains__(self, key):
        return any(key in g for g in [self._module_graph] + self._extra_graphs)

    def __getitem__(self, key):
        for g in [self._module_graph] + self._extra_graphs:
            if key in g:
                return g[key]
        return self._module_graph[key]

    # to be provided by concrete modules --------------------------------------------------------------
    def _generate_outputs(self, output_shapes):
        raise NotImplementedError

    def _build_module_graphs(self):
        raise NotImplementedError

    def _attach_default_inference_algorithms(self):
        raise NotImplementedError

    def set_outputs(self, variables):
        """Attaching the outputs triggers the construction of the inner graphs and default algorithms
        (module.py:104-118)."""
        variables = [variables] if not isinstance(variables, (list, tuple)) else variables
        self.successors = [(n, v) for n, v in zip(self.output_names, variables)]
        self._module_graph, self._extra_graphs = self._build_module_graphs()
        self._attach_default_inference_algorithms()

    # hidden parameters ---------------------------------------------------------------------------------
    def _hidden_parameter_variables(self, excluded=()):
        io = set(v.uuid for _, v in self.inputs) | set(v.uuid for _, v in self.outputs)
        seen, out = set(), []
        for g in [self._module_graph] + self._extra_graphs:
            for var in g.get_parameters(excluded=io | set(excluded), include_inherited=True):
                if var.uuid not in seen:
                    seen.add(var.uuid)
                    out.append(var)
        return out

    @property
    def hidden_parameters(self):
        return [v.uuid for v in self._hidden_parameter_variables()]

    def initialize_hidden_parameters(self, param_dict=None, excluded=None, constants=None):
        """module.py:149-179: declare every inner parameter in the inference parameter store."""
        constants = {} if constants is None else constants
        excluded = set() if excluded is None else set(excluded)
        for var in self._hidden_parameter_variables(excluded | set(constants.keys())):
            param_dict.declare(var, constants)
        return param_dict

    def get_names_from_uuid(self, uuids):
        names = {v.uuid: k for k, v in self.inputs}
        names.update({v.uuid: k for k, v in self.outputs})
        return tuple(sorted(names[u] for u in uuids if u in names))

    # algorithm registries ------------------------------------------------------------------------------
    def attach_log_pdf_algorithms(self, targets, conditionals, algorithm, alg_name=None):
        self._attach_algorithm(self._log_pdf_algorithms, targets, conditionals, algorithm, alg_name)

    def attach_draw_samples_algorithms(self, targets, conditionals, algorithm, alg_name=None):
        self._attach_algorithm(self._draw_samples_algorithms, targets, conditionals, algorithm, alg_name)

    def attach_prediction_algorithms(self, targets, conditionals, algorithm, alg_name=None):
        self._attach_algorithm(self._prediction_algorithms, targets, conditionals, algorithm, alg_name)

    def _attach_algorithm(self, registry, targets, conditionals, algorithm, alg_name):
        from ..inference.inference_alg import InferenceAlgorithm
        targets = tuple(sorted(targets)) if targets is not None else None
        conditionals = tuple(sorted(conditionals)) if conditionals is not None else None
        if alg_name is not None:
            current = self.__dict__.get(alg_name)
            if current is None or isinstance(current, InferenceAlgorithm):
                self.__dict__[alg_name] = algorithm
            else:
                warnings.warn('Something ({}) in this module ({}) is already using the attribute "{}". Skipping '
                              'setting that name to the algorithm.'.format(current, self, alg_name))
                alg_name = None
        entries = registry.setdefault(conditionals, [])
        for i, (t, _, old_name) in enumerate(entries):
            if t == targets:                      # a (targets, conditionals) pair is unique: replace
                if old_name is not None and old_name != alg_name and old_name in self.__dict__:
                    del self.__dict__[old_name]
                entries[i] = (targets, algorithm, alg_name)
                return
        entries.append((targets, algorithm, alg_name))

    def _find_algorithm(self, registry, targets, variables, exact_match=False):
        """module.py:366-391: look up by sorted (targets, conditionals) name tuples."""
        if targets is None:
            target_names = tuple(sorted(self.output_names))
        else:
            target_names = self.get_names_from_uuid(targets)
        cond = self.get_names_from_uuid(variables.keys())
        if exact_match:
            cond = tuple(sorted(set(cond) - set(target_names)))
        if cond in registry:
            want = set(target_names)
            for t, alg, _ in registry[cond]:
                if (exact_match and want == set(t)) or (not exact_match and want <= set(t)):
                    return alg
        raise ModelSpecificationError("The targets-conditionals pattern " + str((target_names, cond)) +
                                      " cannot find a matched inference algorithm.")

    def log_pdf(self, F, variables, targets=None):
        alg = self._find_algorithm(self._log_pdf_algorithms, targets, variables, exact_match=True)
        alg.log_pdf_scaling = self.log_pdf_scaling
        return alg.compute(F, variables)

    def draw_samples(self, F, variables, num_samples=1, targets=None):
        alg = self._find_algorithm(self._draw_samples_algorithms, targets, variables)
        alg.num_samples = num_samples
        alg.target_variables = targets
        return alg.compute(F, variables)

    def predict(self, F, variables, num_samples=1, targets=None):
        alg = self._find_algorithm(self._prediction_algorithms, targets, variables, exact_match=True)
        alg.num_samples = num_samples
        alg.target_variables = targets
        return alg.compute(F, variables)

    def prepare_executor(self, rv_scaling=None):
        """module.py:393-418: collect parameter transformations of the inner graphs, set scalings."""
        var_trans, excluded = {}, set()
        rv_scaling = {} if rv_scaling is None else rv_scaling
        for g in [self._module_graph] + self._extra_graphs:
            for v in g.variables.values():
                if v.type == VariableType.PARAMETER and v.transformation is not None:
                    var_trans[v.uuid] = v.transformation
                if v.type == VariableType.RANDVAR:
                    v.factor.log_pdf_scaling = rv_scaling.get(v.uuid, 1)
        return var_trans, excluded

    def as_json(self):
        d = super(Module, self).as_json()
        d['graphs'] = [g.as_json() for g in [self._module_graph] + self._extra_graphs]
        return d
