"""Generate the committed golden fixtures in tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.pt

Every tensor stored here is an output of reference code (cited per block) on seeded inputs that
are stored next to it, so the fixtures are self-contained and do not depend on RNG reproducibility
across machines.  The reference publishes no golden vectors of its own.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def tiny_llama(dtype=torch.float32):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(
        hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
        num_key_value_heads=4, vocab_size=512, max_position_embeddings=128, tie_word_embeddings=False,
    )
    torch.manual_seed(0)
    m = LlamaForCausalLM(cfg).to(dtype).eval()
    m.config.use_cache = False
    return m


def calib_ids(n=16, t=64, vocab=512):
    # 1024 calibration tokens >= 4x the widest layer (C=256): a well-conditioned Hessian.  With fewer tokens than
    # channels H is rank deficient, Hinv is damp-dominated and codes become chaotic in fp32 summation order.
    g = torch.Generator().manual_seed(1234)
    return [torch.randint(0, vocab, (1, t), generator=g) for _ in range(n)]


def woq_state(model):
    return {k: v.clone() for k, v in model.state_dict().items()
            if any(s in k for s in ("qweight", "qzeros", "scales", "g_idx", "input_scale")) and "bf16_to_fp8" not in k}


def gen_rtn_pack():
    """utility.py:272 quant_tensor, modules.py:321 pack / :413 recover / :594 forward."""
    from neural_compressor.torch.algorithms.weight_only.modules import INCWeightOnlyLinear
    from neural_compressor.torch.algorithms.weight_only.utility import quant_tensor, search_clip

    cases = []
    g = torch.Generator().manual_seed(7)
    specs = [
        # N, K, bits, group, scheme, quantile, full_range, dtype
        (32, 64, 4, 32, "sym", 1.0, False, torch.float32),
        (32, 64, 4, 32, "asym", 1.0, False, torch.float32),
        (64, 256, 4, 128, "sym", 1.0, False, torch.float32),
        (64, 256, 4, 128, "asym", 0.93, False, torch.float32),
        (48, 300, 4, 128, "asym", 1.0, False, torch.float32),   # ragged tail group (utility.py:334-376)
        (48, 300, 4, 128, "sym", 1.0, False, torch.float32),
        (40, 96, 4, -1, "sym", 1.0, True, torch.float32),        # per-channel, full_range (sign flip)
        (40, 96, 8, 32, "sym", 1.0, False, torch.float32),
        (40, 96, 8, 32, "asym", 1.0, False, torch.float32),
        (40, 96, 2, 32, "asym", 1.0, False, torch.float32),
        (24, 128, 3, 32, "sym", 1.0, False, torch.float32),      # n_pack = 10 (32 // 3)
        (32, 128, 4, 32, "sym", 1.0, False, torch.float16),
        (32, 128, 4, 32, "asym", 1.0, False, torch.float16),
        (32, 128, 4, 32, "sym", 0.97, False, torch.bfloat16),
    ]
    for (n, k, bits, gs, scheme, quantile, fr, dt) in specs:
        w = (torch.randn(n, k, generator=g) * 0.05).to(dt)
        if scheme == "asym":
            w[0].abs_()          # an all-positive row: wmin clamps to 0
        w[1].zero_()             # an all-zero row: the (0,0) -> (-1,+1) / amax=1 branch
        wq = w.clone()
        codes, scale, zp = quant_tensor(wq, bits=bits, group_size=gs, scheme=scheme, quantile=quantile,
                                        return_int=True, full_range=fr)
        case = dict(N=n, K=k, bits=bits, group_size=gs, scheme=scheme, quantile=quantile, full_range=fr,
                    W=w, codes=codes.clone(), scale=scale.clone(), zp=None if zp is None else zp.clone())
        fq = quant_tensor(w.clone(), bits=bits, group_size=gs, scheme=scheme, quantile=quantile,
                          return_int=False, full_range=fr)
        case["fake_quant"] = fq.clone()
        eff_g = k if (gs == -1 or k < gs) else gs
        bias = (torch.randn(n, generator=g) * 0.1)
        mod = INCWeightOnlyLinear(k, n, dtype="int", bits=bits, group_size=eff_g, zp=zp is not None,
                                  bias=True, device="cpu")
        mod.pack(codes.clone(), scale.clone(), None if zp is None else zp.clone(), bias)
        case.update(qweight=mod.qweight.clone(), qzeros=mod.qzeros.clone(), scales16=mod.scales.clone(),
                    bias=bias, eff_group=eff_g)
        up = mod.unpack()
        case.update(unpacked_codes=up.get("int_weight").clone(), unpacked_zp=up.get("zp").clone())
        case["recovered"] = mod.recover().clone()
        x = torch.randn(3, k, generator=g)
        case["x"] = x
        case["y"] = mod(x).clone()
        cases.append(case)
    # RTN mse clip search (utility.py:439-480)
    lin = torch.nn.Linear(128, 32)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(32, 128, generator=g) * 0.05)
        lin.weight[:, 5] *= 8
    clip = dict(W=lin.weight.detach().clone(),
                ratio_sym=search_clip(lin, 4, 32, "sym", "int", False),
                ratio_asym=search_clip(lin, 4, 32, "asym", "int", False))
    torch.save(dict(cases=cases, search_clip=clip), os.path.join(OUT, "rtn_pack.pt"))
    print("rtn_pack:", len(cases), "cases")


def gen_config1():
    """BASELINE.json configs[0]: RTN INT4 g128 on nn.Linear(1024,1024) via the public API."""
    from neural_compressor.torch.quantization import RTNConfig, convert, prepare

    torch.manual_seed(0)
    m = torch.nn.Linear(1024, 1024)
    w = m.weight.detach().clone()
    b = m.bias.detach().clone()
    q = convert(prepare(m, RTNConfig(bits=4, group_size=128, use_sym=True, use_layer_wise=False)))
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 1024, generator=g)
    torch.save(dict(W_sum=w.double().sum().item(), W_abs_sum=w.double().abs().sum().item(), bias=b,
                    qweight=q.qweight.clone(), qzeros=q.qzeros.clone(), scales=q.scales.clone(),
                    x=x, y=q(x).clone()),
               os.path.join(OUT, "config1_rtn_linear1024.pt"))
    print("config1: qzeros[0,0] =", hex(q.qzeros[0, 0].item()))


def gen_gptq_layer():
    """gptq.py:1111 add_batch, :1143 fasterquant, utility.py:483 quant_weight_w_scale."""
    from neural_compressor.torch.algorithms.weight_only.gptq import GPTQ
    from neural_compressor.torch.algorithms.weight_only.utility import quant_weight_w_scale

    g = torch.Generator().manual_seed(21)
    N, C, T, S = 48, 256, 64, 6
    lin = torch.nn.Linear(C, N, bias=False)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(N, C, generator=g) * 0.05)
    X = [torch.randn(1, T, C, generator=g) for _ in range(S)]
    for x in X:
        x[..., 3] *= 6.0     # an outlier channel
        x[..., 17] = 0.0     # a dead channel (diag(H) == 0, gptq.py:1189-1191)
    out = dict(W=lin.weight.detach().clone(), X=X, runs=[])
    variants = [
        dict(bits=4, sym=True, group_size=128, blocksize=128, act_order=False, mse=False),
        dict(bits=4, sym=False, group_size=128, blocksize=128, act_order=False, mse=False),
        dict(bits=4, sym=True, group_size=128, blocksize=256, act_order=False, mse=False),   # one block: no lazy GEMM
        dict(bits=4, sym=True, group_size=128, blocksize=2048, act_order=False, mse=False),  # config default, stale find_params
        dict(bits=4, sym=False, group_size=32, blocksize=128, act_order=False, mse=False),
        dict(bits=4, sym=True, group_size=32, blocksize=128, act_order=True, mse=False),
        dict(bits=4, sym=False, group_size=128, blocksize=128, act_order=False, mse=True),
        dict(bits=8, sym=True, group_size=128, blocksize=128, act_order=False, mse=False),
        dict(bits=4, sym=True, group_size=-1, blocksize=128, act_order=False, mse=False),
    ]
    for v in variants:
        gp = GPTQ(lin, lin.weight.data.clone(), "cpu")
        gp.quantizer.configure(dict(dtype="int", bits=v["bits"], sym=v["sym"], group_size=v["group_size"],
                                    mse=v["mse"], perchannel=True, use_double_quant=False, double_quant_sym=False))
        for x in X:
            gp.add_batch(x, None)
        H = gp.H.clone()
        # replay the prologue to record Hinv (gptq.py:1189-1231) without touching the reference
        Hd = H.clone()
        dead = torch.diag(Hd) == 0
        Hd[dead, dead] = 1
        perm = torch.argsort(torch.diag(Hd), descending=True) if v["act_order"] else None
        if perm is not None:
            Hd = Hd[perm][:, perm]
        damp = 0.01 * torch.mean(torch.diag(Hd))
        Hd[torch.arange(C), torch.arange(C)] += damp
        Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
        scale, _, zero, Q = gp.fasterquant(lin.weight.data.clone(), blocksize=v["blocksize"], percdamp=0.01,
                                           groupsize=v["group_size"], act_order=v["act_order"])
        Qe = Q.clone()
        if perm is not None:
            Qe.copy_(Qe[:, perm])
        codes = quant_weight_w_scale(Qe, scale, None, None if v["sym"] else zero, v["group_size"], dtype="int")
        if perm is not None:
            codes.copy_(codes[:, torch.argsort(perm)])
        key = "Hinv_actorder" if v["act_order"] else "Hinv"
        out.setdefault("H", H)          # identical for every variant (same X)
        out.setdefault(key, Hinv)       # depends only on act_order
        assert torch.equal(out["H"], H) and torch.equal(out[key], Hinv)
        out["runs"].append(dict(cfg=v, hinv_key=key, scale=scale.clone(), zero=zero.clone(), Q=Q.clone(),
                                codes=codes.to(torch.int8), perm=perm))
    torch.save(out, os.path.join(OUT, "gptq_layer.pt"))
    print("gptq_layer:", len(out["runs"]), "runs")


def gen_awq_module():
    """awq.py:131-154 stats; search loops restated per module against `quant_tensor`."""
    from neural_compressor.torch.algorithms.weight_only.awq import _get_act_scale, _get_weight_scale
    from neural_compressor.torch.algorithms.weight_only.utility import quant_tensor

    g = torch.Generator().manual_seed(31)
    N, K = 64, 128
    W = torch.randn(N, K, generator=g) * 0.05
    W[:, 9] *= 5
    b = torch.randn(N, generator=g) * 0.1
    X = [torch.randn(1, 24, K, generator=g) * (1 + 3 * (torch.arange(K) % 16 == 0)) for _ in range(5)]
    w_max = _get_weight_scale(W, q_group_size=32)
    x_max = _get_act_scale(X)
    lin = torch.nn.functional.linear
    org = [lin(x, W, b) for x in X]
    hist, cands = [], []
    for i in range(20):
        ratio = i / 20
        s = (x_max.pow(ratio) / w_max.pow(1 - ratio)).clamp(min=1e-4).view(-1)
        s = s / (s.max() * s.min()).sqrt()
        wq = quant_tensor(W.mul(s.view(1, -1)), group_size=32, scheme="asym", full_range=False) / s.view(1, -1)
        hist.append(sum((o - lin(x, wq, b)).float().pow(2).mean().item() for o, x in zip(org, X)))
        cands.append(s)
    chist = []
    for i in range(10):
        ratio = 1 - i / 100
        wq = quant_tensor(W.clone(), group_size=32, scheme="asym", full_range=False, quantile=ratio)
        chist.append(sum((o - lin(x, wq, b)).float().pow(2).mean().item() for o, x in zip(org, X)))
    torch.save(dict(W=W, bias=b, X=X, w_max=w_max, x_max=x_max, scale_hist=hist, scale_cands=torch.stack(cands),
                    clip_hist=chist), os.path.join(OUT, "awq_module.pt"))
    print("awq_module: best scale idx", min(range(20), key=lambda i: hist[i]), "best clip idx",
          min(range(10), key=lambda i: chist[i]))


def gen_e2e():
    """Public API end to end on a tiny random-init Llama: RTN / GPTQ / AWQ (quantize.py:138-325)."""
    from neural_compressor.torch.quantization import AWQConfig, GPTQConfig, RTNConfig, convert, prepare, quantize

    ids = calib_ids()

    def run_fn(model):
        for x in ids:
            model(x)

    base = tiny_llama()
    out = dict(init_state={k: v.clone() for k, v in base.state_dict().items()}, ids=ids)
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(99))
    out["probe"] = probe
    with torch.no_grad():
        out["fp_logits"] = base(probe).logits.clone()

    m = tiny_llama()
    m = convert(prepare(m, RTNConfig(bits=4, group_size=32, use_sym=True, use_layer_wise=False)))
    with torch.no_grad():
        out["rtn"] = dict(state=woq_state(m), logits=m(probe).logits.clone())

    m = tiny_llama()
    m = convert(prepare(m, RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
    with torch.no_grad():
        out["rtn_asym"] = dict(state=woq_state(m), logits=m(probe).logits.clone())

    for tag, kw in (("gptq", dict(use_sym=True, block_size=128)), ("gptq_asym", dict(use_sym=False, block_size=128)),
                    ("gptq_bs2048", dict(use_sym=True))):
        m = tiny_llama()
        cfg = GPTQConfig(bits=4, group_size=32, model_path="/tmp", **kw)
        m = prepare(m, cfg)
        run_fn(m)
        m = convert(m)
        with torch.no_grad():
            out[tag] = dict(state=woq_state(m), logits=m(probe).logits.clone())

    m = tiny_llama()
    cfg = AWQConfig(bits=4, group_size=32, use_sym=False)
    m = quantize(m, cfg, run_fn=run_fn, example_inputs=ids[0])
    with torch.no_grad():
        st = {k: v.clone() for k, v in m.state_dict().items()}
        out["awq"] = dict(state=st, logits=m(probe).logits.clone(),
                          module_types={n: type(x).__name__ for n, x in m.named_modules()})
    torch.save(out, os.path.join(OUT, "e2e_tiny_llama.pt"))
    print("e2e: keys", [k for k in out if k not in ("init_state",)])


OPTION_CASES = [
    # (tag, algo, config kwargs) -- the option surface of the reference's own test matrix (test_rtn.py:106-165,
    # test_gptq.py:106-185, test_awq.py:60-84) on the tiny Llama; RTN must be bit-exact, GPTQ/AWQ within tolerance
    ("rtn_b8_perchannel", "rtn", dict(bits=8, group_size=-1, use_sym=True)),
    ("rtn_b8_asym", "rtn", dict(bits=8, group_size=32, use_sym=False)),
    ("rtn_b4_full_range", "rtn", dict(bits=4, group_size=32, use_sym=True, use_full_range=True)),
    ("rtn_b4_mse_search", "rtn", dict(bits=4, group_size=32, use_sym=False, use_mse_search=True)),
    ("rtn_b3", "rtn", dict(bits=3, group_size=32, use_sym=True)),
    ("rtn_b2_asym", "rtn", dict(bits=2, group_size=64, use_sym=False)),
    ("rtn_lm_head", "rtn", dict(bits=4, group_size=128, use_sym=True, quant_lm_head=True)),
    ("rtn_dtype_int8", "rtn", dict(dtype="int8", group_size=32, use_sym=True)),
    ("gptq_b8", "gptq", dict(bits=8, group_size=32, use_sym=True, block_size=128)),
    ("gptq_perchannel", "gptq", dict(bits=4, group_size=-1, use_sym=True, block_size=128)),
    ("gptq_mse_search", "gptq", dict(bits=4, group_size=32, use_sym=False, use_mse_search=True, block_size=128)),
    ("gptq_act_order", "gptq", dict(bits=4, group_size=32, use_sym=True, act_order=True, block_size=128)),
    ("gptq_true_sequential", "gptq", dict(bits=4, group_size=32, use_sym=True, true_sequential=True, block_size=128)),
    ("gptq_b3", "gptq", dict(bits=3, group_size=32, use_sym=False, block_size=128)),
    ("awq_sym_noclip", "awq", dict(bits=4, group_size=32, use_sym=True, use_auto_clip=False)),
    ("awq_noscale", "awq", dict(bits=4, group_size=32, use_sym=False, use_auto_scale=False)),
]


def gen_options():
    """Option matrix through the reference's public API (one fixture entry per OPTION_CASES row) + a Conv1D model."""
    from neural_compressor.torch.quantization import AWQConfig, GPTQConfig, RTNConfig, convert, prepare, quantize

    ids = calib_ids()
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(99))

    def run_fn(model):
        for x in ids:
            model(x)

    out = dict(cases={})
    for tag, algo, kw in OPTION_CASES:
        m = tiny_llama()
        if algo == "rtn":
            m = convert(prepare(m, RTNConfig(use_layer_wise=False, **kw)))
        elif algo == "gptq":
            m = prepare(m, GPTQConfig(model_path="/tmp", **kw))
            run_fn(m)
            m = convert(m)
        else:
            m = quantize(m, AWQConfig(**kw), run_fn=run_fn, example_inputs=ids[0])
        with torch.no_grad():
            out["cases"][tag] = dict(algo=algo, kw=kw, state=woq_state(m), logits=m(probe).logits.clone())
        print("options:", tag, len(out["cases"][tag]["state"]), "tensors")
    # GPT-2 style model: transformers.Conv1D weights are [in, out] (rtn.py:209-216 transpose handling)
    from transformers import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(0)
    g2 = GPT2LMHeadModel(GPT2Config(n_embd=64, n_layer=2, n_head=2, vocab_size=256, n_positions=64)).eval()
    out["gpt2_init"] = {k: v.clone() for k, v in g2.state_dict().items()}
    p2 = torch.randint(0, 256, (1, 16), generator=torch.Generator().manual_seed(5))
    out["gpt2_probe"] = p2
    g2 = convert(prepare(g2, RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
    with torch.no_grad():
        out["gpt2_rtn"] = dict(state=woq_state(g2), logits=g2(p2).logits.clone(),
                               module_types={n: type(x).__name__ for n, x in g2.named_modules()})
    torch.save(out, os.path.join(OUT, "options_matrix.pt"))


EXTRA_CASES = [
    # f3 leftovers (round 2): double quantisation of the GPTQ scales (gptq.py:1598-1614)
    ("gptq_double_quant", "gptq", dict(bits=4, group_size=32, use_sym=True, block_size=128, use_double_quant=True)),
    ("gptq_double_quant_sym_g64", "gptq", dict(bits=4, group_size=32, use_sym=False, block_size=128, use_double_quant=True,
                                               double_quant_use_sym=True, double_quant_group_size=64)),
]


def gen_options_extra():
    """Round-2 additions to the option matrix, in their own fixture (tests/golden/options_extra.pt), plus the recorded
    behaviour of the reference for `static_groups=True` (it raises IndexError at export)."""
    from neural_compressor.torch.quantization import GPTQConfig, convert, prepare

    ids = calib_ids()
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(99))
    out = dict(cases={})
    for tag, algo, kw in EXTRA_CASES:
        m = tiny_llama()
        m = prepare(m, GPTQConfig(model_path="/tmp", **kw))
        for x in ids:
            m(x)
        m = convert(m)
        with torch.no_grad():
            out["cases"][tag] = dict(algo=algo, kw=kw, state=woq_state(m), logits=m(probe).logits.clone())
        print("options_extra:", tag)
    try:
        m = prepare(tiny_llama(), GPTQConfig(bits=4, group_size=32, static_groups=True, model_path="/tmp"))
        for x in ids:
            m(x)
        convert(m)
        out["static_groups_reference"] = "ok"
    except Exception as ex:  # the reference cannot run its own option (gptq.py:1339-1341 -> utility.py:483-537)
        out["static_groups_reference"] = f"{type(ex).__name__}: {ex}"
    print("static_groups in the reference:", out["static_groups_reference"])
    torch.save(out, os.path.join(OUT, "options_extra.pt"))


def gen_hf_config():
    """AutoGPTQ-style quantization_config the reference derives from a config mapping (save_load.py:1094-1156)."""
    import json

    from neural_compressor.torch.algorithms.weight_only.save_load import change_config_to_hf_format
    from neural_compressor.torch.quantization import AWQConfig, GPTQConfig, RTNConfig

    cases = {}
    g = GPTQConfig(bits=4, group_size=128, use_sym=True, act_order=True, percdamp=0.02)
    cases["gptq"] = dict(args=dict(kind="GPTQConfig", bits=4, group_size=128, use_sym=True, act_order=True, percdamp=0.02),
                         out=change_config_to_hf_format({("model.layers.0.q_proj", "Linear"): g,
                                                         ("lm_head", "Linear"): GPTQConfig(dtype="fp32")}))
    r = RTNConfig(bits=8, group_size=32, use_sym=False)
    cases["rtn"] = dict(args=dict(kind="RTNConfig", bits=8, group_size=32, use_sym=False),
                        out=change_config_to_hf_format({("a", "Linear"): r}))
    a = AWQConfig(bits=4, group_size=64, use_sym=True)
    cases["awq"] = dict(args=dict(kind="AWQConfig", bits=4, group_size=64, use_sym=True),
                        out=change_config_to_hf_format({("a", "Linear"): a}))
    with open(os.path.join(OUT, "hf_quant_config.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("hf_config:", cases)


def gen_awq_repack():
    """AutoAWQ -> optimum repack (utility.py:1432-1459) on random packed tensors: a pure nibble permutation."""
    from neural_compressor.torch.algorithms.weight_only.utility import repack_awq_to_optimum_format

    g = torch.Generator().manual_seed(0)
    K, N, gs = 256, 64, 32
    qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), generator=g, dtype=torch.int64).to(torch.int32)
    qz = torch.randint(-2**31, 2**31 - 1, (K // gs, N // 8), generator=g, dtype=torch.int64).to(torch.int32)
    sc = (torch.rand(K // gs, N, generator=g) * 0.02 + 0.005).half()
    a = repack_awq_to_optimum_format(qw, qz, sc, 4, gs)
    torch.save(dict(awq_qweight=qw, awq_qzeros=qz, awq_scales=sc, group_size=gs, qweight=a[0], qzeros=a[1], scales=a[2]),
               os.path.join(OUT, "awq_repack.pt"))


RTN_DTYPE_CASES = [
    # the reference's own RTN dtype matrix (test_rtn.py:269-300, 302-345): table data types, fp8 casts, double quant
    ("rtn_nf4", dict(dtype="nf4")),
    ("rtn_fp4", dict(dtype="fp4")),
    ("rtn_fp4_e2m1_bnb", dict(dtype="fp4_e2m1_bnb", group_size=64)),
    ("rtn_fp4_e2m1", dict(dtype="fp4_e2m1")),
    ("rtn_nf4_mse", dict(dtype="nf4", use_mse_search=True, group_size=128)),
    ("rtn_fp8_e4m3fn", dict(dtype="fp8_e4m3fn")),
    ("rtn_fp8_e5m2", dict(dtype="fp8_e5m2")),
    ("rtn_int4_dq_asym", dict(dtype="int4", use_double_quant=True, double_quant_bits=6, double_quant_use_sym=False,
                              double_quant_group_size=256)),
    ("rtn_int4_dq_sym", dict(dtype="int4", use_sym=False, use_double_quant=True, double_quant_bits=8,
                             double_quant_use_sym=True, double_quant_group_size=8)),
    ("rtn_nf4_dq", dict(dtype="nf4", use_double_quant=True, double_quant_bits=6, double_quant_use_sym=False,
                        double_quant_group_size=256)),
]


def gen_rtn_dtypes():
    """tests/golden/rtn_dtypes.pt: (1) quantize_4bit under quant_tensor's grouping on seeded tensors (utility.py:121-160,
    272-376); (2) the non-optimum INCWeightOnlyLinear layouts of the reference's module test (test_woq_module.py:10-52:
    bits x compression dtype, plus compression_dim 0 and the no-zero-point case); (3) tiny-llama RTN runs over the
    reference's dtype matrix."""
    import itertools

    from neural_compressor.torch.algorithms.weight_only.modules import INCWeightOnlyLinear
    from neural_compressor.torch.algorithms.weight_only.utility import quant_tensor
    from neural_compressor.torch.quantization import RTNConfig, convert, prepare

    g = torch.Generator().manual_seed(11)
    out = dict(quant=[], rowmajor=[], models={})
    for dtype, T, (n, k, gs), quantile in [("nf4", torch.float32, (32, 128, 32), 1.0), ("nf4", torch.float16, (16, 100, 32), 0.9),
                                           ("nf4", torch.bfloat16, (16, 96, -1), 1.0), ("fp4", torch.float32, (16, 128, 32), 1.0),
                                           ("fp4", torch.float16, (16, 128, 64), 0.95), ("fp4_e2m1", torch.float32, (16, 100, 32), 1.0),
                                           ("fp4_e2m1", torch.bfloat16, (16, 128, 128), 0.805)]:
        w = (torch.randn(n, k, generator=g) * 0.05).to(T)
        w[1, :32] = 0                         # an all-zero group: scale 0, w / scale NaN -> code 0
        codes, scale, _ = quant_tensor(w.clone(), dtype=dtype, group_size=gs, quantile=quantile, return_int=True)
        fake = quant_tensor(w.clone(), dtype=dtype, group_size=gs, quantile=quantile, return_int=False)
        out["quant"].append(dict(dtype=dtype, group_size=gs, quantile=quantile, W=w, codes=codes.to(torch.int8),
                                 scale=scale.float(), fake=fake))
    combos = [(b, cd, 1, "asym") for b, cd in itertools.product([8, 4, 2], [torch.int8, torch.int16, torch.int32, torch.int64])]
    combos += [(4, torch.int32, 0, "asym"), (4, torch.int16, 0, "sym"), (3, torch.int32, 1, "asym"), (4, torch.int32, 1, "sym"),
               (8, torch.int8, 1, "sym"), (2, torch.int64, 0, "asym")]
    for bits, cd, dim, scheme in combos:
        lin = torch.nn.Linear(96, 24)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(24, 96, generator=g) * 0.05)
            lin.bias.copy_(torch.randn(24, generator=g) * 0.01)
        iw, sc, zp = quant_tensor(lin.weight.detach().clone(), dtype="int", bits=bits, return_int=True, group_size=32, scheme=scheme)
        m = INCWeightOnlyLinear(96, 24, dtype="int", bits=bits, group_size=32, zp=zp is not None, bias=True,
                                use_optimum_format=False, compression_dtype=cd, compression_dim=dim, device="cpu")
        m.pack(iw.clone(), sc.clone(), None if zp is None else zp.clone(), lin.bias)
        x = torch.randn(3, 96, generator=g)
        out["rowmajor"].append(dict(bits=bits, compression_dtype=cd, compression_dim=dim, scheme=scheme, int_weight=iw, scale=sc,
                                    zp=zp, bias=lin.bias.detach().clone(), qweight=m.qweight.clone(), scales=m.scales.clone(),
                                    qzeros=m.qzeros.clone() if zp is not None else None, recover=m.recover().clone(), x=x,
                                    y=m(x).clone()))
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(99))
    for tag, kw in RTN_DTYPE_CASES:
        m = tiny_llama()
        m = convert(prepare(m, RTNConfig(**kw)))
        st = woq_state(m)
        if "fp8" in kw["dtype"]:   # the cast weights are exactly representable in the fp8 type: store them in it
            f8 = getattr(torch, kw["dtype"].replace("fp8", "float8"))
            st = {k: v.to(f8) for k, v in m.state_dict().items() if k.endswith("proj.weight")}
        with torch.no_grad():
            out["models"][tag] = dict(kw=kw, state=st, logits=m(probe).logits.clone())
        print("rtn_dtypes:", tag, len(st))
    out["probe"] = probe
    torch.save(out, os.path.join(OUT, "rtn_dtypes.pt"))


def gen_awq_toy():
    """tests/golden/awq_toy.pt: AWQ with discovered absorb layers on a plain nn.Module transformer (tests/toy_models.py),
    the kind of model the reference's GraphTrace can still trace: folding=False (discovered tuples + self-absorbing
    leftovers) and folding=True (folded scales only) -- awq.py:40-95, 364-391."""
    from neural_compressor.torch.algorithms.weight_only.utility import get_absorb_layers
    from neural_compressor.torch.quantization import AWQConfig, quantize

    from tests.toy_models import Toy

    torch.manual_seed(5)
    base = Toy(d=64, n=2, variant=0, vocab=64).eval()
    init = {k: v.clone() for k, v in base.state_dict().items()}
    g = torch.Generator().manual_seed(17)
    ids = [torch.randint(0, 64, (1, 16), generator=g) for _ in range(8)]
    probe = torch.randint(0, 64, (1, 16), generator=g)
    out = dict(init_state=init, ids=ids, probe=probe, cases={})
    absorb, no_absorb = get_absorb_layers(base, ids[0], supported_layers=["Linear"])
    out["absorb_to_layer"], out["no_absorb_layers"] = absorb, no_absorb

    def run_fn(model):
        for x in ids:
            model(x)

    for tag, kw in [("folding_false", dict(folding=False)), ("folding_true", dict(folding=True)),
                    ("folding_true_sym", dict(folding=True, use_sym=True, group_size=64))]:
        m = Toy(d=64, n=2, variant=0, vocab=64).eval()
        m.load_state_dict(init)
        cfg = AWQConfig(bits=4, group_size=kw.pop("group_size", 32), use_sym=kw.pop("use_sym", False), **kw)
        m = quantize(m, cfg, run_fn=run_fn, example_inputs=ids[0])
        with torch.no_grad():
            st = {k: v.clone() for k, v in m.state_dict().items() if "bf16_to_fp8" not in k}
            out["cases"][tag] = dict(state=st, logits=m(probe).clone())
        print("awq_toy:", tag, sorted({k.rsplit(".", 1)[-1] for k in st}), len(st))
    torch.save(out, os.path.join(OUT, "awq_toy.pt"))


def gen_sq_transform():
    """tests/golden/sq_transform.pt: the reference's SmoothQuant smoothing transform (smooth_quant/utility.py:
    Calibration :840-953, cal_scale :605-626, TorchSmoothQuant._cal_scales :2122-2156, _scale_layer_weight :1968-1992,
    _absorb_scales :1994-2061, SQLinearWrapper :2559-2662) on the CPU with IPEX stubbed (oracle/ref_loader.py), plus the
    QDQ simulation helpers (quant_dequant_w_v1 :652-690, quant_dequant_x_v1 :726-755) on seeded tensors.  Models: the
    traceable toy transformer (scale sharing / folding through GraphTrace) and the tiny llama (GraphTrace fails under
    transformers 5: every Linear gets its own scale, folding finds nothing)."""
    from oracle.ref_loader import load_smooth_quant_utility

    SQ = load_smooth_quant_utility()
    from tests.toy_models import Toy

    out = dict(models={}, helpers={})
    g = torch.Generator().manual_seed(23)
    w = torch.randn(48, 96, generator=g) * 0.05
    x = torch.randn(64, 96, generator=g) * 2.0
    lin = torch.nn.Linear(96, 48, bias=False)
    with torch.no_grad():
        lin.weight.copy_(w)
    mn, mx = x.min(0)[0], x.max(0)[0]
    s = SQ.cal_scale(torch.max(mn.abs(), mx.abs()), [w], 0.5)
    out["helpers"] = dict(w=w, x=x, cal_scale=s, cal_scale_a08=SQ.cal_scale(torch.max(mn.abs(), mx.abs()), [w, w * 2], 0.8),
                          qdq_w=SQ.quant_dequant_w_v1(lin).clone(), qdq_x=SQ.quant_dequant_x_v1(x.clone()).clone(),
                          qdq_x_minmax=SQ.quant_dequant_x_v1(x.clone(), min_x=torch.tensor(-3.0), max_x=torch.tensor(2.5)).clone())

    def run(tag, build, ids, probe, folding, scale_sharing=True, alpha=0.5):
        m = build()

        def q_func(model):
            for t in ids:
                model(t)

        sq = SQ.TorchSmoothQuant(m, dataloader=None, example_inputs=ids[0], q_func=q_func, scale_sharing=scale_sharing)
        m = sq.transform(alpha=alpha, folding=folding, calib_iter=len(ids), op_types=[torch.nn.Linear])
        st = {}
        for n, mod in m.named_modules():
            if type(mod).__name__ == "SQLinearWrapper":
                st[n] = dict(input_scale=mod.input_scale.clone(), scale=mod.scale.clone(), zero_point=mod.zero_point.clone(),
                             weight=mod.sq_linear.weight.detach().clone())
        with torch.no_grad():
            out["models"][tag] = dict(folding=folding, scale_sharing=scale_sharing, alpha=alpha, wrappers=st,
                                      absorb_to_layer=dict(sq.absorb_to_layer or {}),
                                      input_mins={k: v.clone() for k, v in sq.input_mins.items()},
                                      input_maxes={k: v.clone() for k, v in sq.input_maxes.items()},
                                      state={k: v.clone() for k, v in m.state_dict().items()} if folding else None,
                                      logits=(m(probe)[0] if isinstance(m(probe), tuple) else getattr(m(probe), "logits", m(probe))).clone())
        print("sq_transform:", tag, len(st), "wrappers; absorb_to_layer", {k: v for k, v in list((sq.absorb_to_layer or {}).items())[:3]})

    awq = torch.load(os.path.join(OUT, "awq_toy.pt"))

    def toy():
        m = Toy(d=64, n=2, variant=0, vocab=64).eval()
        m.load_state_dict(awq["init_state"])
        return m

    for tag, kw in [("toy_insert_mul", dict(folding=False)), ("toy_insert_mul_noshare", dict(folding=False, scale_sharing=False)),
                    ("toy_folding", dict(folding=True)), ("toy_insert_mul_a08", dict(folding=False, alpha=0.8))]:
        run(tag, toy, awq["ids"], awq["probe"], **kw)
    ids = calib_ids(n=8, t=32)
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(99))
    run("llama_insert_mul", tiny_llama, ids, probe, folding=False)
    out["llama_ids"], out["llama_probe"] = ids, probe
    # alpha="auto": AutoAlpha's model-wise search (:1232-1893); the tuned per-group alphas
    out["auto"] = {}
    for tag, build, data, args in [
            ("toy_mean", toy, awq["ids"], dict(init_alpha=0.5, alpha_min=0.0, alpha_max=1.0, alpha_step=0.1, shared_criterion="mean", n_samples=8)),
            ("toy_max", toy, awq["ids"], dict(init_alpha=0.5, alpha_min=0.3, alpha_max=0.7, alpha_step=0.05, shared_criterion="max", n_samples=6)),
            ("llama_mean", tiny_llama, ids, dict(init_alpha=0.5, alpha_min=0.0, alpha_max=1.0, alpha_step=0.1, shared_criterion="mean", n_samples=8))]:
        def q_func(model, data=data):
            for t in data:
                model(t)

        sq = SQ.TorchSmoothQuant(build(), dataloader=None, example_inputs=data[0], q_func=q_func)
        sq.transform(alpha="auto", folding=False, calib_iter=8, op_types=[torch.nn.Linear], auto_alpha_args=dict(args))
        out["auto"][tag] = dict(args=args, alpha=dict(sq.alpha))
        print("sq_transform: auto", tag, list(sq.alpha.items())[:3])
    torch.save(out, os.path.join(OUT, "sq_transform.pt"))


def gen_gptq_hybrid():
    """tests/golden/gptq_hybrid.pt: `hybrid_order=True` (gptq.py:1203-1209, 1320-1328, 1389-1474): columns sorted by
    diag(H) inside their group, groups by their largest entry; group parameters put back in storage order afterwards."""
    from neural_compressor.torch.quantization import GPTQConfig, convert, prepare

    ids = calib_ids()
    probe = torch.randint(0, 512, (1, 16), generator=torch.Generator().manual_seed(99))
    out = dict(cases={}, probe=probe)
    for tag, kw in [("hybrid_sym", dict(bits=4, group_size=32, use_sym=True, block_size=128, hybrid_order=True)),
                    ("hybrid_asym_g64", dict(bits=4, group_size=64, use_sym=False, block_size=128, hybrid_order=True))]:
        m = prepare(tiny_llama(), GPTQConfig(model_path="/tmp", **kw))
        for x in ids:
            m(x)
        m = convert(m)
        with torch.no_grad():
            out["cases"][tag] = dict(kw=kw, state=woq_state(m), logits=m(probe).logits.clone())
        print("gptq_hybrid:", tag)
    torch.save(out, os.path.join(OUT, "gptq_hybrid.pt"))


def family_models():
    """Tiny random-init models of other architectures than Llama: OPT (biased linears, LayerNorm, ReLU MLP), GPT-J
    (parallel attention / MLP reading one LayerNorm) and GPT-2 (transformers.Conv1D weights stored [in, out])."""
    from transformers import GPT2Config, GPT2LMHeadModel, GPTJConfig, GPTJForCausalLM, OPTConfig, OPTForCausalLM

    def build(name):
        torch.manual_seed(0)
        if name == "opt":
            m = OPTForCausalLM(OPTConfig(hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=256,
                                         max_position_embeddings=128, word_embed_proj_dim=64))
        elif name == "gptj":
            m = GPTJForCausalLM(GPTJConfig(n_embd=64, n_layer=2, n_head=4, n_positions=128, vocab_size=256, rotary_dim=8))
        else:
            m = GPT2LMHeadModel(GPT2Config(n_embd=64, n_layer=2, n_head=4, vocab_size=256, n_positions=128))
        m.eval()
        m.config.use_cache = False
        return m

    return build


def gen_families():
    """tests/golden/families.pt: GPTQ and AWQ on tiny OPT / GPT-J / GPT-2 models (the reference's own tests run on a tiny
    GPT-J): calibration capture with each architecture's block signature, biased linears, Conv1D transposes."""
    from neural_compressor.torch.quantization import AWQConfig, GPTQConfig, convert, prepare, quantize

    build = family_models()
    ids = calib_ids(n=8, t=32, vocab=256)
    probe = torch.randint(0, 256, (1, 16), generator=torch.Generator().manual_seed(99))
    out = dict(ids=ids, probe=probe, init={}, cases={})

    def run_fn(model):
        for x in ids:
            model(x)

    for name in ("opt", "gptj", "gpt2"):
        out["init"][name] = {k: v.clone() for k, v in build(name).state_dict().items()}
        try:
            m = prepare(build(name), GPTQConfig(bits=4, group_size=32, use_sym=False, block_size=128, model_path="/tmp"))
            run_fn(m)
            m = convert(m)
            with torch.no_grad():
                out["cases"][f"gptq_{name}"] = dict(state=woq_state(m), logits=m(probe).logits.clone())
            print("families: gptq", name, len(out["cases"][f"gptq_{name}"]["state"]))
        except Exception as ex:   # GPT-2: the reference's own export trips over the Conv1D layout (gptq.py:796-813)
            out["cases"][f"gptq_{name}"] = dict(reference_error=f"{type(ex).__name__}: {ex}")
            print("families: gptq", name, "REFERENCE FAILS:", out["cases"][f"gptq_{name}"]["reference_error"][:100])
        if name != "gpt2":     # AWQ walks nn.Linear modules
            m = quantize(build(name), AWQConfig(bits=4, group_size=32, use_sym=False), run_fn=run_fn, example_inputs=ids[0])
            with torch.no_grad():
                out["cases"][f"awq_{name}"] = dict(state=woq_state(m), logits=m(probe).logits.clone())
            print("families: awq", name, len(out["cases"][f"awq_{name}"]["state"]))
    torch.save(out, os.path.join(OUT, "families.pt"))


GENERATORS_EXTRA = {"options_extra": gen_options_extra, "awq_repack": gen_awq_repack, "rtn_dtypes": gen_rtn_dtypes,
                    "awq_toy": gen_awq_toy, "sq_transform": gen_sq_transform, "gptq_hybrid": gen_gptq_hybrid,
                    "families": gen_families}

if __name__ == "__main__":
    load_reference()
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["rtn", "config1", "gptq", "awq", "e2e", "hf_config", "options"]
    with torch.no_grad():
        if "rtn" in which:
            gen_rtn_pack()
        if "config1" in which:
            gen_config1()
        if "gptq" in which:
            gen_gptq_layer()
        if "awq" in which:
            gen_awq_module()
        if "e2e" in which:
            gen_e2e()
        if "hf_config" in which:
            gen_hf_config()
        if "options" in which:
            gen_options()
        if "options_extra" in which:
            gen_options_extra()
        if "awq_repack" in which:
            gen_awq_repack()
        if "rtn_dtypes" in which:
            gen_rtn_dtypes()
        if "awq_toy" in which:
            gen_awq_toy()
        if "sq_transform" in which:
            gen_sq_transform()
        if "gptq_hybrid" in which:
            gen_gptq_hybrid()
        if "families" in which:
            gen_families()
