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
        """Restores parameter values saved by `save` into this (already constructed, same-topology)
        inference.  Variables are matched by graph position + name (UUIDs differ between processes)."""
        with zipfile.ZipFile(zip_filename, 'r') as zf:
            version = json.loads(zf.read(FILENAMES['version_file']))
            if version['serialization_version'] != SERIALIZATION_VERSION:
                raise SerializationError("Serialization version of saved inference and running code are not "
                                         "the same.")
            graphs = json.loads(zf.read(FILENAMES['graphs']))
            arrays = dict(np.load(io.BytesIO(zf.read(FILENAMES['mxnet_params']))))
        uuid_map = {}
        for saved, cur in zip(graphs, self._graphs):
            cur_json = cur.as_json()
            for a, b in zip(saved['components'], cur_json['components']):
                if a['type'] != b['type'] or a['name'] != b['name']:
                    raise SerializationError("saved graph does not match the current graph at component %s / %s"
                                             % (a, b))
                uuid_map[a['uuid']] = b['uuid']
            for sm, cm in zip([c for c in saved['components'] if 'graphs' in c],
                              [c for c in cur_json['components'] if 'graphs' in c]):
                for sg, cg in zip(sm['graphs'], cm['graphs']):
                    for a, b in zip(sg['components'], cg['components']):
                        uuid_map[a['uuid']] = b['uuid']
        self._loaded_arrays = {uuid_map.get(k, k): v for k, v in arrays.items()}
        if self._initialized:
            self._apply_loaded()

    def _apply_loaded(self):
        loaded = getattr(self, '_loaded_arrays', None)
        if not loaded:
            return
        for uuid, p in self.params.param_dict.items():
            if uuid in loaded and tuple(loaded[uuid].shape) == tuple(p.tensor.shape):
                p.set_data(torch.as_tensor(loaded[uuid]))


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
