"""Quantizer contract.  Reference: neural_compressor/torch/algorithms/base_algorithm.py:25-126."""
from abc import ABC, abstractmethod
from typing import Any, Optional

import torch

from ..utils import Mode


class Quantizer(ABC):
    def __init__(self, quant_config: Optional[Any] = None):
        self.quant_config = quant_config

    @abstractmethod
    def prepare(self, model: torch.nn.Module, *args: Any, **kwargs: Any):
        raise NotImplementedError(f"{self.__class__.__name__} doesn't implement `prepare` function. ")

    @abstractmethod
    def convert(self, model: torch.nn.Module, *args: Any, **kwargs: Any):
        raise NotImplementedError(f"{self.__class__.__name__} doesn't implement `convert` function. ")

    def quantize(self, model: torch.nn.Module, *args: Any, **kwargs: Any):
        """prepare -> run_fn(model, *run_args) -> convert (base_algorithm.py:79-101)."""
        model = self.prepare(model, *args, **kwargs)
        run_fn = kwargs.get("run_fn", None)
        if run_fn is not None:
            run_args = kwargs.get("run_args", None)
            if run_args:
                run_fn(model, *run_args)
            else:
                run_fn(model)
        return self.convert(model, *args, **kwargs)

    def execute(self, model: torch.nn.Module, mode, *args: Any, **kwargs: Any):
        # compared by VALUE: when the entries are registered into the reference's own registry the
        # reference passes ITS `Mode` enum (common/utils/constants.py:55-62), a different class with the same values
        mode = Mode(getattr(mode, "value", mode))
        if mode == Mode.PREPARE:
            return self.prepare(model, *args, **kwargs)
        if mode == Mode.CONVERT:
            return self.convert(model, *args, **kwargs)
        if mode == Mode.QUANTIZE:
            return self.quantize(model, *args, **kwargs)
        raise ValueError(f"unknown mode {mode}")
