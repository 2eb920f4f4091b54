"""Normalising-flow interface (mirror of mae/modules/flows/flow.py:9-76).  Flow networks stay in PyTorch;
the hot path only consumes their outputs ``(z_normal, logdet)``."""
from typing import Dict, Tuple

import torch
import torch.nn as nn


class Flow(nn.Module):
    _registry: Dict[str, type] = dict()

    def input_size(self) -> Tuple:
        raise NotImplementedError

    def forward(self, x: torch.Tensor, init=False, init_scale=1.0) -> Tuple[torch.Tensor, torch.Tensor]:
        """x -> (y, logdet) with log p(y) = log p(x) + logdet."""
        raise NotImplementedError

    def backward(self, y: torch.Tensor, init=False, init_scale=1.0) -> Tuple[torch.Tensor, torch.Tensor]:
        """y -> (x, logdet) with log p(y) = log p(x) + logdet."""
        raise NotImplementedError

    @classmethod
    def register(cls, name: str):
        Flow._registry[name] = cls

    @classmethod
    def by_name(cls, name: str):
        return Flow._registry[name]

    @classmethod
    def from_params(cls, params: Dict):
        raise NotImplementedError
