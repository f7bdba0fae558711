import pytest
import torch

requires_gpu = pytest.mark.gpu


def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda")
