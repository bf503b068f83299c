"""The small calibration helpers of weight_only/utility.py, with the expectations of the reference's own unit tests
(test/torch/quantization/weight_only/test_woq_utils.py, test/torch/algorithms/weight_only/test_woq_utility.py)."""
import pytest
import torch

from neural_compressor_b200.utils.utility import (CapturedDataloader, convert_dtype_str2torch, forward_wrapper,
                                                    get_example_input, model_forward, move_input_to_device)


def test_move_input_to_device():
    d = move_input_to_device({"a": torch.tensor([1, 2, 3]), "b": torch.tensor([4, 5, 6])})
    assert all(v.device.type == "cpu" for v in d.values())
    lst = move_input_to_device([torch.tensor([1]), (torch.tensor([2]), "x")])
    assert isinstance(lst, list) and isinstance(lst[1], tuple) and lst[1][1] == "x"
    assert move_input_to_device("string") == "string"


def test_forward_wrapper():
    class Double(torch.nn.Module):
        def forward(self, x=None, **kw):
            return (x if x is not None else kw["input_ids"]) * 2

    t = torch.tensor([1, 2, 3])
    assert torch.all(forward_wrapper(Double(), t) == t * 2)
    assert torch.all(forward_wrapper(Double(), {"input_ids": t}) == t * 2)

    class Add:
        def to(self, device):
            return self

        def __call__(self, x, y):
            return x + y

    assert torch.all(forward_wrapper(Add(), [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6])]) == torch.tensor([5, 7, 9]))

    class Boom(Add):
        def __call__(self, x):
            raise ValueError("Mock model exception")

    with pytest.raises(ValueError):
        forward_wrapper(Boom(), [torch.tensor([1]), torch.tensor([2])])


def test_model_forward_and_example_input():
    calls = []

    class Model:
        def to(self, device):
            return self

        def __call__(self, x):
            calls.append(x)
            return x * 2

    pairs = [(torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6])), (torch.tensor([7, 8, 9]), torch.tensor([10, 11, 12])),
             (torch.tensor([13, 14, 15]), torch.tensor([16, 17, 18]))]
    model_forward(Model(), pairs, 1, torch.device("cpu"))
    assert len(calls) == 1 and torch.equal(calls[0], torch.tensor([1, 2, 3]))
    calls.clear()
    model_forward(Model(), [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6])], -1, torch.device("cpu"))
    assert len(calls) == 2
    assert torch.equal(get_example_input(pairs, i=1), torch.tensor([7, 8, 9]))
    assert torch.equal(get_example_input(pairs, i=0), torch.tensor([1, 2, 3]))
    assert torch.equal(get_example_input([torch.tensor([5]), torch.tensor([6])], i=1), torch.tensor([6]))


def test_captured_dataloader():
    a, b = torch.tensor([1]), torch.tensor([2])
    out = list(CapturedDataloader([(a,), (), (a, b), (a,)], [{}, {"input_ids": b}, {}, {"mask": b}]))
    assert out[0] is a and out[1] == {"input_ids": b} and out[2] == (a, b) and out[3] == ((a,), {"mask": b})


def test_convert_dtype_str2torch():
    assert convert_dtype_str2torch("int8") == torch.int8
    for s in ("fp32", "float32", "auto"):
        assert convert_dtype_str2torch(s) == torch.float
    assert convert_dtype_str2torch("fp16") == torch.float16 and convert_dtype_str2torch("float16") == torch.float16
    assert convert_dtype_str2torch("bf16") == torch.bfloat16 and convert_dtype_str2torch("bfloat16") == torch.bfloat16
    assert convert_dtype_str2torch(torch.float16) == torch.float16 and convert_dtype_str2torch(None) is None
    with pytest.raises(AssertionError):
        convert_dtype_str2torch("int16")
