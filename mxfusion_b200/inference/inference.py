"""Inference drivers (mxfusion/inference/inference.py:31-358)."""
import io
import json
import warnings
import zipfile

import numpy as np
import torch

from .inference_parameters import InferenceParameters, discover_shape_constants
from ..common.config import get_default_dtype, get_default_device, torch_dtype
from ..common.exceptions import InferenceError, SerializationError
from ..models import Model, Posterior

SERIALIZATION_VERSION = '1.0'
FILENAMES = {'graphs': 'graphs.json', 'mxnet_params': 'mxnet_parameters.npz',
             'mxnet_constants': 'mxnet_constants.npz', 'variable_constants': 'variable_constants.json',
             'configuration': 'configuration.json', 'version_file': 'version.json'}
DEFAULT_ZIP = 'inference.zip'


class Inference(object):
    """`Inference(alg, constants, hybridize, dtype, context)`; `context` is a torch.device."""

    def __init__(self, inference_algorithm, constants=None, hybridize=False, dtype=None, context=None):
        self.dtype = dtype if dtype is not None else get_default_dtype()
        self.mxnet_context = torch.device(context) if context is not None else get_default_device()
        self._hybridize = hybridize
        self._graphs = inference_algorithm.graphs
        self._inference_algorithm = inference_algorithm
        self.params = InferenceParameters(constants=constants, dtype=self.dtype, context=self.mxnet_context)
        self._initialized = False

    def print_params(self):
        out = ""
        for uuid, p in self.params.param_dict.items():
            owner = [(g, g[uuid]) for g in self._graphs if uuid in g]
            if not owner:
                continue
            g, var = owner[0]
            kind = "Model" if isinstance(g, Model) else "Posterior" if isinstance(g, Posterior) else "FactorGraph"
            out += "{} in {}({}) : {} \n\n".format(var, kind, g._uuid[:5], p.data())
        return out

    @property
    def observed_variables(self):
        return self._inference_algorithm.observed_variables

    @property
    def observed_variable_UUIDs(self):
        return self._inference_algorithm.observed_variable_UUIDs

    @property
    def observed_variable_names(self):
        return self._inference_algorithm.observed_variable_names

    @property
    def graphs(self):
        return self._graphs

    @property
    def inference_algorithm(self):
        return self._inference_algorithm

    def create_executor(self):
        return self._inference_algorithm.create_executor(data_def=self.observed_variable_UUIDs, params=self.params,
                                                         var_ties=self.params.var_ties)

    def _initialize_params(self):
        self.params.initialize_params(self._graphs, self.observed_variable_UUIDs)

    def initialize(self, **kw):
        """Shapes (tuples) or data arrays keyed by variable name (inference.py:126-156)."""
        if self._initialized:
            warnings.warn("Trying to initialize the inference twice, skipping.")
            return
        data = [kw[v] for v in self.observed_variable_names]
        if len(data) > 0:
            if isinstance(data[0], (tuple, list)):
                shapes = {i: tuple(d) for i, d in zip(self.observed_variable_UUIDs, data)}
            elif isinstance(data[0], (torch.Tensor, np.ndarray)):
                shapes = {i: tuple(d.shape) for i, d in zip(self.observed_variable_UUIDs, data)}
            else:
                raise InferenceError("Keywords not of type array or tuple/list for shapes passed into "
                                     "initialization.")
            self.params.update_constants(discover_shape_constants(shapes, self._graphs))
        self._initialize_params()
        self._initialized = True
        self._apply_loaded()

    def _to_device(self, d):
        t = torch.as_tensor(d)
        if t.is_floating_point():
            t = t.to(torch_dtype(self.dtype))
        return t.to(self.mxnet_context)

    def run(self, **kwargs):
        data = [self._to_device(kwargs[v]) for v in self.observed_variable_names]
        self.initialize(**kwargs)
        executor = self.create_executor()
        return executor(None, *data)

    # checkpoint (inference.py:179-310): same archive layout -- one zip with six members ------------------
    def get_serializable(self):
        return {'observed': self.observed_variable_UUIDs}

    def save(self, zip_filename=DEFAULT_ZIP):
        params, tensor_consts, other_consts = self.params.get_serializable()
        payload = [(FILENAMES['graphs'], json.dumps([g.as_json() for g in self._graphs])),
                   (FILENAMES['variable_constants'], json.dumps(other_consts)),
                   (FILENAMES['configuration'], json.dumps(self.get_serializable())),
                   (FILENAMES['version_file'], json.dumps({"serialization_version": SERIALIZATION_VERSION}))]
        with zipfile.ZipFile(zip_filename, 'w', zipfile.ZIP_DEFLATED) as zf:
            for name, text in payload:
                zf.writestr(name, text)
            for name, arrays in ((FILENAMES['mxnet_params'], params), (FILENAMES['mxnet_constants'], tensor_consts)):
                buf = io.BytesIO()
                np.savez(buf, **arrays)
                zf.writestr(name, buf.getvalue())

    def load(self, zip_filename=DEFAULT_ZIP):
        """Restores parameter values saved by `save` into this (already constructed) inference.  UUIDs differ between
        processes, so the saved graphs are reconciled with the current ones (factor_graph.py:479-590 does this by name
        and topology): within a graph, components are matched by (type, name) -- independent of the order in which the
        model was written down -- and components that share a (type, name) pair by their order of appearance.  Any
        component, parameter or shape that does not match raises SerializationError; nothing is skipped silently.
        The values are applied ONCE: right away if the inference is initialised, else at the end of `initialize`."""
        with zipfile.ZipFile(zip_filename, 'r') as zf:
            version = json.loads(zf.read(FILENAMES['version_file']))
            if version['serialization_version'] != SERIALIZATION_VERSION:
                raise SerializationError("Serialization version of saved inference and running code are not "
                                         "the same.")
            graphs = json.loads(zf.read(FILENAMES['graphs']))
            arrays = dict(np.load(io.BytesIO(zf.read(FILENAMES['mxnet_params']))))
        if len(graphs) != len(self._graphs):
            raise SerializationError("saved inference holds %d graphs, this one %d" % (len(graphs), len(self._graphs)))
        uuid_map = {}
        for saved, cur in zip(graphs, self._graphs):
            _reconcile_components(saved['components'], cur.as_json()['components'], uuid_map,
                                  saved.get('name') or saved.get('class'))
        unknown = [k for k in arrays if k not in uuid_map]
        if unknown:
            raise SerializationError("saved parameters %s belong to no component of the saved graphs" % unknown[:3])
        self._loaded_arrays = {uuid_map[k]: v for k, v in arrays.items()}
        if self._initialized:
            self._apply_loaded()

    def _apply_loaded(self):
        """Copies the loaded values into the parameters and forgets them, so that later `run` calls continue from the
        trained state instead of being reset to the checkpoint."""
        loaded = getattr(self, '_loaded_arrays', None)
        if not loaded:
            return
        self._loaded_arrays = None
        missing = [u for u in loaded if u not in self.params.param_dict]
        if missing:
            raise SerializationError("%d saved parameters have no counterpart in this inference (first: %s)"
                                     % (len(missing), missing[0]))
        for uuid, value in loaded.items():
            p = self.params.param_dict[uuid]
            if tuple(value.shape) != tuple(p.tensor.shape):
                raise SerializationError("saved parameter %s has shape %s, the current one %s"
                                         % (uuid, tuple(value.shape), tuple(p.tensor.shape)))
            p.set_data(torch.as_tensor(value))


def _reconcile_components(saved, cur, uuid_map, where):
    """saved / cur: `as_json()['components']` lists of one graph.  Fills uuid_map[saved uuid] = current uuid."""
    from collections import Counter

    def key(c):
        return (c['type'], c['name'])
    if Counter(key(c) for c in saved) != Counter(key(c) for c in cur):
        only_saved = sorted(str(k) for k in (Counter(key(c) for c in saved) - Counter(key(c) for c in cur)))
        only_cur = sorted(str(k) for k in (Counter(key(c) for c in cur) - Counter(key(c) for c in saved)))
        raise SerializationError("saved graph %r does not match the current graph: only saved %s, only current %s"
                                 % (where, only_saved[:4], only_cur[:4]))
    by_key, taken = {}, Counter()
    for c in cur:
        by_key.setdefault(key(c), []).append(c)
    for a in saved:
        b = by_key[key(a)][taken[key(a)]]
        taken[key(a)] += 1
        uuid_map[a['uuid']] = b['uuid']
        ga, gb = a.get('graphs', []), b.get('graphs', [])
        if len(ga) != len(gb):
            raise SerializationError("module %s holds %d graphs in the archive, %d now" % (key(a), len(ga), len(gb)))
        for sg, cg in zip(ga, gb):
            _reconcile_components(sg['components'], cg['components'], uuid_map, '%s/%s' % (where, a['name']))


class TransferInference(Inference):
    """Run an algorithm with the parameters learned by a previous inference (inference.py:313-358)."""

    def __init__(self, inference_algorithm, infr_params, var_tie=None, constants=None, hybridize=False, dtype=None,
                 context=None):
        self._var_tie = var_tie if var_tie is not None else {}
        self._inherited_params = infr_params if isinstance(infr_params, list) else [infr_params]
        super(TransferInference, self).__init__(inference_algorithm=inference_algorithm, constants=constants,
                                                hybridize=hybridize, dtype=dtype, context=context)

    def _initialize_params(self):
        self.params.initialize_with_carryover_params(self._graphs, self.observed_variable_UUIDs, self._var_tie,
                                                     self._inherited_params)
