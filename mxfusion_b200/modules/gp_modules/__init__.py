from .gp_regression import (GPRegression, GPRegressionLogPdf, GPRegressionMeanVariancePrediction,  # noqa: F401
                            GPRegressionSamplingPrediction)
from .svgp_regression import (SVGPRegression, SVGPRegressionLogPdf, SVGPRegressionMeanVariancePrediction,  # noqa: F401
                              SVGPRegressionSamplingPrediction)
from .sparsegp_regression import (SparseGPRegression, SparseGPRegressionLogPdf,  # noqa: F401
                                  SparseGPRegressionMeanVariancePrediction, SparseGPRegressionSamplingPrediction)
from .deep_gp import DeepGPRegression, DeepGPLogPdf, DeepGPMeanVariancePrediction  # noqa: F401
