"""mxfusion/inference/grad_loop.py:19-45."""
from abc import ABC, abstractmethod


class GradLoop(ABC):
    @abstractmethod
    def run(self, infr_executor, data, param_dict, ctx, optimizer='adam', learning_rate=1e-3, max_iter=1000,
            verbose=False):
        pass
