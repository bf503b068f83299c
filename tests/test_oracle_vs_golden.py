"""Pin oracle/woq_oracle.py against fixtures produced by the unmodified reference (CPU only)."""
import pytest
import torch

from oracle import woq_oracle as O


def _case_id(c):
    return f"N{c['N']}K{c['K']}b{c['bits']}g{c['group_size']}{c['scheme']}q{c['quantile']}{'fr' if c['full_range'] else ''}{str(c['W'].dtype)[6:]}"


def test_rtn_quantize_matches_reference(golden_rtn):
    for c in golden_rtn["cases"]:
        q, s, z = O.rtn_quantize(c["W"], c["bits"], c["group_size"], c["scheme"], c["quantile"], c["full_range"])
        assert torch.equal(q, c["codes"]), _case_id(c)
        assert torch.equal(s, c["scale"]), _case_id(c)
        assert (z is None) == (c["zp"] is None)
        if z is not None:
            assert torch.equal(z, c["zp"]), _case_id(c)
        fq = O.rtn_fake_quant(c["W"], c["bits"], c["group_size"], c["scheme"], c["quantile"], c["full_range"])
        assert torch.equal(fq, c["fake_quant"]), _case_id(c)


def test_pack_unpack_recover_forward_match_reference(golden_rtn):
    for c in golden_rtn["cases"]:
        qw, qz, sc = O.pack_optimum(c["codes"], c["scale"], c["zp"], c["bits"], c["eff_group"])
        assert torch.equal(qw, c["qweight"]), _case_id(c)
        assert torch.equal(qz, c["qzeros"]), _case_id(c)
        assert torch.equal(sc, c["scales16"]), _case_id(c)
        codes, zp, _ = O.unpack_optimum(qw, qz, sc, c["bits"], c["K"], c["N"])
        assert torch.equal(codes.long(), c["unpacked_codes"].long()), _case_id(c)
        assert torch.equal(zp.long(), c["unpacked_zp"].long()), _case_id(c)
        rec = O.recover_fp16(qw, qz, sc, c["bits"], c["eff_group"], c["K"], c["N"])
        assert rec.dtype == torch.float16 and torch.equal(rec, c["recovered"]), _case_id(c)
        y = O.woq_linear_forward(c["x"], qw, qz, sc, c["bias"].half(), c["bits"], c["eff_group"], c["K"], c["N"])
        assert torch.equal(y, c["y"]), _case_id(c)


def test_rtn_search_clip_matches_reference(golden_rtn):
    sc = golden_rtn["search_clip"]
    assert O.rtn_search_clip(sc["W"], 4, 32, "sym") == sc["ratio_sym"]
    assert O.rtn_search_clip(sc["W"], 4, 32, "asym") == sc["ratio_asym"]


def test_config1_linear1024(golden_config1):
    """BASELINE.json configs[0]."""
    g = golden_config1
    torch.manual_seed(0)
    m = torch.nn.Linear(1024, 1024)
    w = m.weight.detach()
    if abs(w.double().sum().item() - g["W_sum"]) > 1e-9:
        pytest.skip("torch RNG stream differs from the fixture's")
    q, s, z = O.rtn_quantize(w, 4, 128, "sym")
    qw, qz, sc = O.pack_optimum(q, s, z, 4, 128)
    assert torch.equal(qw, g["qweight"]) and torch.equal(qz, g["qzeros"]) and torch.equal(sc, g["scales"])
    assert (qz == 0x77777777).all()
    y = O.woq_linear_forward(g["x"], qw, qz, sc, g["bias"].half(), 4, 128, 1024, 1024)
    assert torch.equal(y, g["y"])


def _gptq_run(golden_gptq, run, inject_hinv):
    v = run["cfg"]
    W = golden_gptq["W"]
    N, C = W.shape
    lay = O.GPTQLayerOracle(N, C, bits=v["bits"], sym=v["sym"], mse=v["mse"])
    for x in golden_gptq["X"]:
        lay.add_batch(x)
    if inject_hinv:
        Wp = W.clone()
        Wp[:, torch.diag(lay.H) == 0] = 0
        perm = run["perm"]
        if perm is not None:
            Wp = Wp[:, perm]
        res = lay.fasterquant(Wp, v["blocksize"], 0.01, v["group_size"], hinv=golden_gptq[run["hinv_key"]])
        if perm is not None:
            res["Q"] = res["Q"][:, torch.argsort(perm)]
            res["perm"] = perm
    else:
        res = lay.fasterquant(W, v["blocksize"], 0.01, v["group_size"], act_order=v["act_order"])
    return lay, res


@pytest.mark.parametrize("inject_hinv", [True, False])
def test_gptq_layer_matches_reference(golden_gptq, inject_hinv):
    for run in golden_gptq["runs"]:
        v = run["cfg"]
        lay, res = _gptq_run(golden_gptq, run, inject_hinv)
        assert torch.equal(lay.H, golden_gptq["H"])
        if not inject_hinv:
            assert torch.equal(res["hinv"], golden_gptq[run["hinv_key"]]), v
        assert torch.equal(res["scale"], run["scale"]), v
        assert torch.equal(res["zero"], run["zero"]), v
        assert torch.equal(res["Q"], run["Q"]), v
        codes = O.GPTQLayerOracle.export_codes(res["Q"], res["scale"], res["zero"], v["group_size"], v["sym"], res["perm"])
        assert torch.equal(codes, run["codes"].to(torch.int32)), v


def test_awq_stats_and_search(golden_awq):
    g = golden_awq
    assert torch.equal(O.awq_weight_scale(g["W"], 32), g["w_max"])
    assert torch.equal(O.awq_act_scale(g["X"]), g["x_max"])
    s, r, hist = O.awq_search_scale_module(g["W"], g["bias"], g["X"], 32, "asym")
    assert hist == g["scale_hist"]
    best = min(range(20), key=lambda i: g["scale_hist"][i])
    assert torch.equal(s, g["scale_cands"][best]) and r == best / 20
    ratio, chist = O.awq_search_clip_module(g["W"], g["bias"], g["X"], 32, "asym")
    assert chist == g["clip_hist"]
    assert ratio == 1 - min(range(10), key=lambda i: g["clip_hist"][i]) / 100


def test_oracle_reproduces_reference_rtn_option_matrix(golden_e2e, golden_options):
    """Pins the oracle's RTN + pack restatement on the option surface: for every RTN case of the option matrix
    (bits 2/3/4/8, per-channel, asym, full_range, mse_search, quantised lm_head) the oracle applied to the initial
    weights must give exactly the packed tensors the live reference produced through its public API."""
    init = golden_e2e["init_state"]
    checked = 0
    for tag, case in golden_options["cases"].items():
        if case["algo"] != "rtn":
            continue
        kw = case["kw"]
        bits = int(kw["dtype"].lstrip("int")) if "dtype" in kw else kw["bits"]
        scheme = "sym" if kw["use_sym"] else "asym"
        for key, ref in case["state"].items():
            if not key.endswith(".qweight"):
                continue
            name = key[: -len(".qweight")]
            W = init[name + ".weight"].float()
            g = kw["group_size"]
            quantile = O.rtn_search_clip(W, bits, g, scheme, kw.get("use_full_range", False)) if kw.get("use_mse_search") else 1.0
            q, s, z = O.rtn_quantize(W, bits, g, scheme, quantile, kw.get("use_full_range", False))
            geff = W.shape[1] if g == -1 else g
            qw, qz, sc = O.pack_optimum(q, s, z, bits, geff)
            assert torch.equal(qw, ref), (tag, key)
            assert torch.equal(qz, case["state"][name + ".qzeros"]), (tag, name)
            assert torch.equal(sc, case["state"][name + ".scales"]), (tag, name)
            checked += 1
    assert checked >= 8 * 14
