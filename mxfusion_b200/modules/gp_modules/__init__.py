from .gp_regression import GPRegression, GPRegressionLogPdf, GPRegressionMeanVariancePrediction  # noqa: F401
from .svgp_regression import SVGPRegression, SVGPRegressionLogPdf, SVGPRegressionMeanVariancePrediction  # noqa: F401
from .sparsegp_regression import (SparseGPRegression, SparseGPRegressionLogPdf,  # noqa: F401
                                  SparseGPRegressionMeanVariancePrediction, SparseGPRegressionSamplingPrediction)
