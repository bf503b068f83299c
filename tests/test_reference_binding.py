"""INTEGRATION.md §2 executed: our algorithm entries are registered into the UNMODIFIED reference's own registry with its
`register_algo`, and the reference's own `prepare()/convert()` + `RTNConfig/GPTQConfig` objects drive them.

* `not gpu`: the dispatch reaches our entry with the reference's config objects and stops at the first CUDA op with our
  loud `B200WOQError` (there is no CPU path) -- the seam, the attribute names and the error convention are exercised.
* `gpu`: the same binding runs for real (reference imported from oracle/_ref on the GPU box) and must produce exactly
  what our mirrored API produces.
The reference is imported from /root/reference (build container) or from the vendored copy oracle/_ref."""
import pytest
import torch


def _reference():
    from oracle import ref_loader

    if not (ref_loader.reference_available() or ref_loader.vendored_available()):
        pytest.skip("reference neither mounted nor vendored (oracle/build_ref.py)")
    ref_loader.load_reference()
    import neural_compressor.torch.quantization as RQ
    from neural_compressor.torch.utils import algos_mapping, register_algo

    return RQ, algos_mapping, register_algo


def _bind(register_algo):
    from neural_compressor_b200.quantization.algorithm_entry import awq_quantize_entry, gptq_entry, rtn_entry

    register_algo("rtn")(rtn_entry)
    register_algo("gptq")(gptq_entry)
    register_algo("awq")(awq_quantize_entry)
    return rtn_entry


def test_register_algo_binding_dispatches_into_b200_entries():
    RQ, algos_mapping, register_algo = _reference()
    saved = dict(algos_mapping)
    try:
        entry = _bind(register_algo)
        assert algos_mapping["rtn"] is entry
        m = torch.nn.Sequential(torch.nn.Linear(64, 32))
        if torch.cuda.is_available():
            pytest.skip("covered by the gpu variant")
        from neural_compressor_b200._lib import B200WOQError

        with pytest.raises(B200WOQError):   # reference prepare/convert -> our rtn_entry -> first CUDA op: loud failure
            RQ.convert(RQ.prepare(m, RQ.RTNConfig(bits=4, group_size=32, use_layer_wise=False)))
    finally:
        algos_mapping.clear()
        algos_mapping.update(saved)


@pytest.mark.gpu
def test_register_algo_binding_runs_on_the_b200(golden_rtn):
    RQ, algos_mapping, register_algo = _reference()
    import neural_compressor_b200.quantization as Q
    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear

    saved = dict(algos_mapping)
    try:
        _bind(register_algo)
        torch.manual_seed(0)
        m1 = torch.nn.Sequential(torch.nn.Linear(256, 128)).cuda()
        m2 = torch.nn.Sequential(torch.nn.Linear(256, 128)).cuda()
        m2.load_state_dict(m1.state_dict())
        # the reference's prepare/convert with the reference's RTNConfig, our kernels underneath
        q1 = RQ.convert(RQ.prepare(m1, RQ.RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
        q2 = Q.convert(Q.prepare(m2, Q.RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
        assert isinstance(q1[0], B200WeightOnlyLinear)
        for k in ("qweight", "qzeros", "scales"):
            assert torch.equal(getattr(q1[0], k), getattr(q2[0], k)), k
    finally:
        algos_mapping.clear()
        algos_mapping.update(saved)
