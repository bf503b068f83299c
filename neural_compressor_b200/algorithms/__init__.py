from .base_algorithm import Quantizer
