from .inference_alg import InferenceAlgorithm, SamplingAlgorithm, ObjectiveBlock  # noqa: F401
from .variational import VariationalInference, StochasticVariationalInference  # noqa: F401
from .map import MAP  # noqa: F401
from .meanfield import create_Gaussian_meanfield, create_Gaussian_process  # noqa: F401
from .inference_parameters import InferenceParameters  # noqa: F401
from .grad_loop import GradLoop  # noqa: F401
from .batch_loop import BatchInferenceLoop  # noqa: F401
from .minibatch_loop import MinibatchInferenceLoop, RolloverBatchSampler  # noqa: F401
from .inference import Inference, TransferInference  # noqa: F401
from .grad_based_inference import GradBasedInference  # noqa: F401
from .prediction import ModulePredictionAlgorithm  # noqa: F401
from .forward_sampling import (ForwardSamplingAlgorithm, ForwardSampling,  # noqa: F401
                               VariationalPosteriorForwardSampling, VariationalPosteriorForwardSamplingAlgorithm)
