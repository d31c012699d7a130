import contextlib
import torch


@contextlib.contextmanager
def record(train_mode=True):
    with torch.enable_grad():
        yield


@contextlib.contextmanager
def pause(train_mode=False):
    with torch.no_grad():
        yield


def is_recording():
    return torch.is_grad_enabled()
