"""HF-style entry of the weight-only path, mirroring `neural_compressor.transformers`
(__init__.py:15-26): `AutoModelForCausalLM.from_pretrained(path, quantization_config=RtnConfig(...))`."""
from .utils import AwqConfig, GPTQConfig, QuantizationMethod, RtnConfig  # noqa: F401
from .models import AutoModelForCausalLM, convert_to_quantized_model  # noqa: F401
